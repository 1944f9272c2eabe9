#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clustering.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/t.log 2>&1; echo "clustering tests exit $?"; tail -4 gpurun_out/t.log
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/tp.log 2>&1; echo "pipeline tests exit $?"; tail -3 gpurun_out/tp.log
for fz in 1 0; do
  UOC_FPS_PRUNED=$fz timeout 600 python bench.py --steps 30 --warmup 3 --depth 3 --no-cpu-baseline > gpurun_out/bench_fps$fz.json 2> gpurun_out/bench_fps$fz.err; echo "bench fps_pruned=$fz exit $?"
  python -c "
import json; j=json.load(open('gpurun_out/bench_fps$fz.json')); print(round(j['value'],1), round(j['e2e']['value'],1), 'serial', round(j['serial']['value'],1), j['stages_ms'], j['roofline']['frac'], j['gpu_launches'])"
done
