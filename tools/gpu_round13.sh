#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_backbone.py -m gpu -q --timeout 300 -p no:cacheprovider -k "conv_matches_torch and (halo or tc_cluster1)" > gpurun_out/test_conv.log 2>&1; echo "conv tests exit $?"; grep -E "passed|failed|FAILED" gpurun_out/test_conv.log | head -40
timeout 600 python -m pytest tests/test_gpu_clustering.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/test_gpu_clustering.log 2>&1; echo "clustering exit $?"; tail -2 gpurun_out/test_gpu_clustering.log
for h in 0 1; do
UOC_CONV_HALO=$h timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --depth 3 > gpurun_out/bench_h$h.json 2> gpurun_out/bench_h$h.err
python -c "
import json; j=json.load(open('gpurun_out/bench_h$h.json')); print('halo $h', round(j['value'],1), round(j['e2e']['value'],1), 'serial', round(j['serial']['value'],1), j['stages_ms'])"
tail -2 gpurun_out/bench_h$h.err
done
