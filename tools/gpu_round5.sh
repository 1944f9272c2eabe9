#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
./tools/mma_bench > gpurun_out/mma_bench.txt 2>&1; cat gpurun_out/mma_bench.txt
timeout 600 python -m pytest tests/test_gpu_clustering.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/test_gpu_clustering.log 2>&1; echo "clustering exit $?"; tail -2 gpurun_out/test_gpu_clustering.log
for cfg in "1 128" "1 256" "2 256" "4 128" "2 128"; do
set -- $cfg
UOC_CONV_CLUSTER=$1 UOC_CONV_MAX_BLOCK_N=$2 timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c$1_n$2.json 2> gpurun_out/bench_c$1_n$2.err
python -c "
import json; j=json.load(open('gpurun_out/bench_c$1_n$2.json')); print('cluster $1 n $2', round(j['value'],1), j['stages_ms'])"
done
