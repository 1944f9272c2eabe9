#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
python tools/fps_trace.py 2>&1 | grep "cta   0\|cta  73" | tee gpurun_out/fps_trace.txt
timeout 600 python -m pytest tests/test_gpu_clustering.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/test_gpu_clustering.log 2>&1; echo "clustering exit $?"; tail -2 gpurun_out/test_gpu_clustering.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
python -c "
import json; j=json.load(open('gpurun_out/bench_a.json')); print(round(j['value'],1), j['stages_ms'], j['roofline']['frac'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
