#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/test_gpu_pipeline.log 2>&1; echo "pipeline exit $?"; tail -15 gpurun_out/test_gpu_pipeline.log
timeout 900 python bench.py --steps 50 --warmup 3 --depth 3 --no-cpu-baseline > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench exit $?"
python -c "
import json; j=json.load(open('gpurun_out/bench_g.json')); print(round(j['value'],1), round(j['e2e']['value'],1), 'serial', round(j['serial']['value'],1), j['stages_ms'], 'host', j['host_enqueue_ms_per_step'], j['gpu_launches'])"
tail -3 gpurun_out/bench_g.err
