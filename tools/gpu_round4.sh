#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for f in test_gpu_clustering test_gpu_backbone test_gpu_pipeline; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/$f.log 2>&1; echo "$f exit $?"; tail -4 gpurun_out/$f.log
done
for cl in 4 8 1; do
UOC_CONV_CLUSTER=$cl timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cl$cl.json 2> gpurun_out/bench_cl$cl.err; echo "bench cl$cl exit $?"
python -c "
import json; j=json.load(open('gpurun_out/bench_cl$cl.json')); print('cluster $cl', j['value'], j['ms_per_step'], j['stages_ms'], j['roofline']['frac'], j['e2e']['value'])"
done
UOC_CONV_CLUSTER=4 UOC_CONV_MAX_BLOCK_N=128 timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cl4_n128.json 2>&1
python -c "
import json; j=json.load(open('gpurun_out/bench_cl4_n128.json')); print('cluster 4 n128', j['value'], j['stages_ms'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
