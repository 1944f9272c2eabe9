#!/bin/bash
# full GPU validation: every -m gpu test file in its own process, smoke, bench (N=1)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for f in test_gpu_clustering test_gpu_backbone test_gpu_pipeline test_gpu_input_prep; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/$f.log 2>&1; echo "$f exit $?"; tail -2 gpurun_out/$f.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 50 --warmup 3 --depth 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"
python -c "
import json; j=json.load(open('gpurun_out/bench_final.json')); print(round(j['value'],1), round(j['e2e']['value'],1), 'serial', round(j['serial']['value'],1), j['stages_ms'], 'host', j.get('host_loop_ms_per_step'), j['clocks'], j['cpu_baseline'], j['gpu_launches'])"
