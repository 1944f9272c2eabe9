#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_backbone.py -m gpu -q --timeout 300 -p no:cacheprovider -k "conv_matches_torch and halo_small" > gpurun_out/test_conv.log 2>&1; echo "conv tests exit $?"; tail -2 gpurun_out/test_conv.log
UOC_CONV_HALO=1 UOC_CONV_HALO_SMALL=1 UOC_CONV_SUB=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_halo4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --depth 1 > gpurun_out/bench_under_ncu.log 2>&1
for cfg in "0 0" "1 1"; do
set -- $cfg
UOC_CONV_HALO=$1 UOC_CONV_HALO_SMALL=$2 UOC_CONV_SUB=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --depth 3 > gpurun_out/bench_hs$1.json 2> gpurun_out/bench_hs$1.err
python -c "
import json; j=json.load(open('gpurun_out/bench_hs$1.json')); print('halo small $1', round(j['value'],1), 'serial', round(j['serial']['value'],1), j['stages_ms'])"
done
