#!/bin/bash
# session-2 call W: streaming bf16-screen sampler (fps5_kernel): parity, config 5 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clustering.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/t_cluster.log 2>&1; echo "cluster exit $?"; tail -12 gpurun_out/t_cluster.log
timeout 300 python tools/bench_configs.py > gpurun_out/bench_configs.log 2>&1; echo "configs exit $?"; cut -c1-330 gpurun_out/bench_configs.log | tail -2
