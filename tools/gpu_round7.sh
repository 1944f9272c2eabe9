#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
./tools/mma_bench > gpurun_out/mma_bench3.txt 2>&1; grep -E "accumulators|TS 128x64x16 \(A|SS 128x128x16  " gpurun_out/mma_bench3.txt
timeout 600 python -m pytest tests/test_gpu_clustering.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/test_gpu_clustering.log 2>&1; echo "clustering exit $?"; tail -3 gpurun_out/test_gpu_clustering.log
for v in 0 1; do
UOC_FPS_VARIANT=$v UOC_CONV_CLUSTER=1 UOC_CONV_MAX_BLOCK_N=128 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fpsv$v.json 2> gpurun_out/bench_fpsv$v.err
python -c "
import json; j=json.load(open('gpurun_out/bench_fpsv$v.json')); print('fps variant $v', round(j['value'],1), j['stages_ms'])"
done
