#!/bin/bash
# session-2 call B: TMEM / MUFU microbenchmark, batched-frames A/B
mkdir -p gpurun_out
timeout 120 ./tools/tmem_bench > gpurun_out/tmem_bench.txt 2>&1; echo "tmem_bench exit $?"; cat gpurun_out/tmem_bench.txt
timeout 600 python tools/batch_ab.py > gpurun_out/batch_ab.log 2>&1; echo "batch_ab exit $?"; tail -8 gpurun_out/batch_ab.log
