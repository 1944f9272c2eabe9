#!/bin/bash
# session-2 call T: two-stage plumbing kernels (refine.cu): parity, stage profile, config 3
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_pipeline.log 2>&1; echo "pipeline exit $?"; tail -25 gpurun_out/t_pipeline.log
timeout 300 python tools/two_stage_profile.py 2>&1 | tail -10
timeout 300 python tools/bench_configs.py 2>&1 | head -1 | cut -c1-300
