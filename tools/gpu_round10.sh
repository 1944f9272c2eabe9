#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for f in test_gpu_clustering test_gpu_pipeline; do
timeout 600 python -m pytest tests/$f.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/$f.log 2>&1; echo "$f exit $?"; tail -3 gpurun_out/$f.log
done
for d in 2 3 1; do
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --depth $d > gpurun_out/bench_d$d.json 2> gpurun_out/bench_d$d.err
python -c "
import json; j=json.load(open('gpurun_out/bench_d$d.json')); print('depth $d', round(j['value'],1), round(j['e2e']['value'],1), 'serial', round(j['serial']['value'],1), j['stages_ms'])"
tail -2 gpurun_out/bench_d$d.err
done
UOC_FPS_HEAVY=1 timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --depth 2 > gpurun_out/bench_heavy.json 2> gpurun_out/bench_heavy.err
python -c "
import json; j=json.load(open('gpurun_out/bench_heavy.json')); print('heavy fps depth 2', round(j['value'],1), round(j['e2e']['value'],1), 'serial', round(j['serial']['value'],1), j['stages_ms'])"
