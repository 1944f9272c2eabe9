#!/bin/bash
# session-2 call E: split GEMM1 / GEMM2 issue in the persistent loop, frames_per_slot pipeline, bench batch 1 vs 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_clustering.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/t_cluster.log 2>&1; echo "cluster exit $?"; tail -5 gpurun_out/t_cluster.log
timeout 300 python tools/loop_ab.py > gpurun_out/loop_ab.log 2>&1; echo "loop_ab exit $?"; tail -12 gpurun_out/loop_ab.log
UOC_LOOP_POLY=8 timeout 120 python tools/loop_trace.py > /dev/null 2> gpurun_out/loop_trace_split_poly8.txt; grep "update [45]" gpurun_out/loop_trace_split_poly8.txt | grep -v "P-1" | tail -4
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_pipeline.log 2>&1; echo "pipeline exit $?"; tail -15 gpurun_out/t_pipeline.log
for b in 1 2; do
  timeout 600 python bench.py --steps 30 --warmup 3 --depth 3 --batch $b --no-cpu-baseline > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err; echo "bench batch=$b exit $?"; tail -3 gpurun_out/bench_b$b.err
  python -c "
import json; j=json.load(open('gpurun_out/bench_b$b.json')); print(round(j['value'],1), round(j['e2e']['value'],1), round(j['e2e_raw_inputs']['value'],1), 'serial', round(j['serial']['value'],1), j['stages_ms'], j['roofline']['frac'])"
done
