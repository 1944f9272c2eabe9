#!/bin/bash
# round 2, session 4: where does the CTA-pair convolution lose its time?  (hang hunt with per-test time-outs first)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_backbone.py -m gpu -x -v --timeout 40 -k "test_conv_matches_torch and pair" 2>&1 | tail -40 > gpurun_out/r2s4_tests.log
tail -25 gpurun_out/r2s4_tests.log
export UOC_CONV_LAYERS_ONLY="l3 3x3 256 d2,l4 3x3 512 d4,l2 3x3 128"
UOC_CONV_LAYERS_VARIANTS="pair;UOC_CONV_TRACE=1;UOC_CONV_DEBUG=1;UOC_CONV_DEBUG=2;UOC_CONV_DEBUG=3;UOC_CONV_DEBUG=4;UOC_CONV_DEBUG=8;UOC_CONV_PAIR=0" timeout 300 python tools/conv_layers.py 2 8 > gpurun_out/r2s4_conv_debug.txt 2>&1
cat gpurun_out/r2s4_conv_debug.txt
