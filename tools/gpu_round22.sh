#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_backbone.py -m gpu -q --timeout 300 -p no:cacheprovider -k "not conv_matches_torch" > gpurun_out/test_gpu_backbone.log 2>&1; echo "backbone exit $?"; tail -12 gpurun_out/test_gpu_backbone.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 30 --warmup 3 --depth 3 --no-cpu-baseline > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err; echo "bench exit $?"
python -c "
import json; j=json.load(open('gpurun_out/bench_s.json')); print(round(j['value'],1), round(j['e2e']['value'],1), 'serial', round(j['serial']['value'],1), j['stages_ms'])"
