#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_input_prep.py tests/test_gpu_pipeline.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/t.log
timeout 600 python bench.py --steps 30 --warmup 3 --depth 3 --no-cpu-baseline > gpurun_out/bench_prep.json 2> gpurun_out/bench_prep.err; echo "bench exit $?"; tail -3 gpurun_out/bench_prep.err
python -c "
import json; j=json.load(open('gpurun_out/bench_prep.json')); print(round(j['value'],1), 'e2e', round(j['e2e']['value'],1), 'raw', round(j['e2e_raw_inputs']['value'],1), 'serial', round(j['serial']['value'],1), j['stages_ms'], j['roofline']['frac'], j['roofline']['traffic'])"
