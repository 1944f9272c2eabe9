#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_clustering.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/test_gpu_clustering.log 2>&1; echo "clustering exit $?"; tail -3 gpurun_out/test_gpu_clustering.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench exit $?"
python -c "
import json; j=json.load(open('gpurun_out/bench_a.json')); print(j['value'], j['ms_per_step'], j['stages_ms'], j['roofline']['frac'], j['e2e']['value'])"
UOC_FPS_SMEM_KB=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fps_nosmem.json 2>&1
python -c "
import json; j=json.load(open('gpurun_out/bench_fps_nosmem.json')); print('no smem', j['stages_ms'])"
UOC_FPS_V1=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fps_v1.json 2>&1
python -c "
import json; j=json.load(open('gpurun_out/bench_fps_v1.json')); print('v1', j['stages_ms'])"
tail -3 gpurun_out/bench_a.err
