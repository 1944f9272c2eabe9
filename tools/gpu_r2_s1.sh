#!/bin/bash
# round 2, session 1: full GPU test suite + the new bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2s1_tests.log
tail -5 gpurun_out/r2s1_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2s1_bench.json 2> gpurun_out/r2s1_bench.err
tail -c 3000 gpurun_out/r2s1_bench.err
head -c 6000 gpurun_out/r2s1_bench.json
