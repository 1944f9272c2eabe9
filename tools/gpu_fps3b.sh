#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clustering.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "select_seeds or golden or batched or cfg or config" > gpurun_out/t.log 2>&1; echo "clustering tests exit $?"; tail -3 gpurun_out/t.log
python tools/fps3_trace.py 0 70 2>&1 | grep "fps3 trace" | grep -E "mean|pass (0|1|2|3|5|10|40|50|90):" | cut -c1-200
timeout 600 python bench.py --steps 30 --warmup 3 --depth 3 --no-cpu-baseline > gpurun_out/bench_fps1.json 2> gpurun_out/bench_fps1.err; echo "bench exit $?"
python -c "
import json; j=json.load(open('gpurun_out/bench_fps1.json')); print(round(j['value'],1), round(j['e2e']['value'],1), 'serial', round(j['serial']['value'],1), j['stages_ms'], j['roofline']['frac'], j['gpu_launches'])"
