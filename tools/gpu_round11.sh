#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for cfg in "96 2" "96 3" "48 3" "0 3" "200 3"; do
set -- $cfg
UOC_FPS_HEAVY=1 UOC_FPS_SMEM_KB=$1 timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --depth $2 > gpurun_out/bench_s$1_d$2.json 2> gpurun_out/bench_s$1_d$2.err
python -c "
import json; j=json.load(open('gpurun_out/bench_s$1_d$2.json')); print('heavy smem $1 depth $2', round(j['value'],1), round(j['e2e']['value'],1), 'serial', round(j['serial']['value'],1), j['stages_ms'])"
done
