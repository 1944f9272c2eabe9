#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clustering.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "select_seeds" > gpurun_out/t.log 2>&1; echo "select_seeds tests exit $?"; tail -5 gpurun_out/t.log
timeout 120 python tools/fps_tc_trace.py 0 70 2>&1 | grep "trace" | cut -c1-260
timeout 600 python bench.py --steps 30 --warmup 3 --depth 3 --no-cpu-baseline > gpurun_out/bench_fps4.json 2> gpurun_out/bench_fps4.err; echo "bench exit $?"
python -c "
import json; j=json.load(open('gpurun_out/bench_fps4.json')); print(round(j['value'],1), round(j['e2e']['value'],1), 'serial', round(j['serial']['value'],1), j['stages_ms'], j['roofline']['frac'], j['gpu_launches'])"
