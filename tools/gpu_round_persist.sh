#!/bin/bash
# persistent mean-shift loop: parity tests + A/B bench
mkdir -p gpurun_out
for f in test_gpu_clustering test_gpu_pipeline; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/$f.log 2>&1; echo "$f exit $?"; tail -3 gpurun_out/$f.log
done
for pz in 1 0; do
  UOC_LOOP_PERSISTENT=$pz timeout 600 python bench.py --steps 30 --warmup 3 --depth 3 --no-cpu-baseline > gpurun_out/bench_persist$pz.json 2> gpurun_out/bench_persist$pz.err; echo "bench persist=$pz exit $?"
  python -c "
import json; j=json.load(open('gpurun_out/bench_persist$pz.json')); print(round(j['value'],1), round(j['e2e']['value'],1), 'serial', round(j['serial']['value'],1), j['stages_ms'], j['roofline']['frac'], j['gpu_launches'])"
done
