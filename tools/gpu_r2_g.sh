#!/bin/bash
# session-2 call G: bench sweep over frames per step / steps in flight
mkdir -p gpurun_out
for cfg in "1 3" "2 2" "2 3" "3 2" "4 2" "4 3"; do
  set -- $cfg
  timeout 600 python bench.py --steps 24 --warmup 3 --batch $1 --depth $2 --no-cpu-baseline > gpurun_out/bench_b$1_d$2.json 2> gpurun_out/bench_b$1_d$2.err; echo "bench batch=$1 depth=$2 exit $?"
  python -c "
import json; j=json.load(open('gpurun_out/bench_b$1_d$2.json')); print('  value', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), 'raw', round(j['e2e_raw_inputs']['value'],1), 'serial', round(j['serial']['value'],1), 'host ms/step', j['host_loop_ms_per_step'])"
done
