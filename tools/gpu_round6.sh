#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
./tools/mma_bench > gpurun_out/mma_bench2.txt 2>&1; cat gpurun_out/mma_bench2.txt
for v in 0 1 2 3; do
UOC_FPS_VARIANT=$v UOC_CONV_CLUSTER=1 UOC_CONV_MAX_BLOCK_N=256 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fpsv$v.json 2> gpurun_out/bench_fpsv$v.err
python -c "
import json; j=json.load(open('gpurun_out/bench_fpsv$v.json')); print('fps variant $v', round(j['value'],1), j['stages_ms'])"
done
UOC_FPS_VARIANT=1 UOC_FPS_SMEM_KB=0 UOC_CONV_CLUSTER=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fpsv1_nosmem.json 2>&1
python -c "
import json; j=json.load(open('gpurun_out/bench_fpsv1_nosmem.json')); print('fps variant 1 nosmem', round(j['value'],1), j['stages_ms'])"
