#!/bin/bash
# round 2, session 2: the tests that did not run in session 1 + per-layer convolution timings (A/B of the CTA-pair kernel)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_flip_rate.py tests/test_gpu_input_prep.py tests/test_gpu_pipeline.py tests/test_gpu_baselines.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r2s2_tests.log
tail -30 gpurun_out/r2s2_tests.log
python tools/conv_layers.py 2 8 > gpurun_out/r2s2_conv_layers.txt 2>&1
cat gpurun_out/r2s2_conv_layers.txt
