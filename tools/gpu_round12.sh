#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for f in test_gpu_clustering test_gpu_pipeline; do
timeout 600 python -m pytest tests/$f.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/$f.log 2>&1; echo "$f exit $?"; tail -4 gpurun_out/$f.log
done
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --depth 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
python -c "
import json; j=json.load(open('gpurun_out/bench_a.json')); print('depth 3', round(j['value'],1), round(j['e2e']['value'],1), 'serial', round(j['serial']['value'],1), j['stages_ms'])"
tail -3 gpurun_out/bench_a.err
