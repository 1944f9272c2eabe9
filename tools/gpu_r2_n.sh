#!/bin/bash
# session-2 call N: all GPU tests, smoke, bench (defaults), configs 3 and 5
mkdir -p gpurun_out
for f in test_gpu_clustering test_gpu_backbone test_gpu_pipeline test_gpu_input_prep; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/$f.log 2>&1; echo "$f exit $?"; tail -2 gpurun_out/$f.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/bench_configs.py > gpurun_out/bench_configs.log 2>&1; echo "configs exit $?"; tail -3 gpurun_out/bench_configs.log
timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err; echo "bench exit $?"; tail -3 gpurun_out/bench_s2.err
python -c "
import json; j=json.load(open('gpurun_out/bench_s2.json')); print(round(j['value'],1), round(j['e2e']['value'],1), round(j['e2e_raw_inputs']['value'],1), 'serial', round(j['serial']['value'],1), j['stages_ms'], j['roofline']['frac'], j['clocks'], j['cpu_baseline'], j['gpu_launches'])"
