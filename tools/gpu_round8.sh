#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
./tools/mma_bench > gpurun_out/mma_bench4.txt 2>&1; grep -E "^SS|^TS" gpurun_out/mma_bench4.txt
python tools/fps_trace.py 2>&1 | tee gpurun_out/fps_trace.txt
for f in test_gpu_clustering test_gpu_backbone; do
timeout 600 python -m pytest tests/$f.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/$f.log 2>&1; echo "$f exit $?"; tail -2 gpurun_out/$f.log
done
for cfg in "1 128" "1 256"; do
set -- $cfg
UOC_CONV_CLUSTER=$1 UOC_CONV_MAX_BLOCK_N=$2 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c$1_n$2.json 2> gpurun_out/bench_c$1_n$2.err
python -c "
import json; j=json.load(open('gpurun_out/bench_c$1_n$2.json')); print('cluster $1 n $2', round(j['value'],1), j['stages_ms'], j['roofline']['frac'])"
done
