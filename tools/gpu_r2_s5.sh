#!/bin/bash
# round 2, session 5: pair kernel with the TMA-store epilogue: parity (per-test time-outs), per-layer A/B, trace, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backbone.py -m gpu -x -v --timeout 60 2>&1 | tail -30 > gpurun_out/r2s5_tests.log
tail -12 gpurun_out/r2s5_tests.log
UOC_CONV_LAYERS_VARIANTS="pair;UOC_CONV_PAIR=0" timeout 300 python tools/conv_layers.py 2 8 > gpurun_out/r2s5_conv_layers.txt 2>&1
cat gpurun_out/r2s5_conv_layers.txt
UOC_CONV_LAYERS_ONLY="l3 3x3 256 d2,l1 3x3 64,l4 1x1 down" UOC_CONV_LAYERS_VARIANTS="UOC_CONV_TRACE=1" timeout 300 python tools/conv_layers.py 8 2>&1 | grep "conv trace" | awk 'NR%7==1' > gpurun_out/r2s5_conv_trace.txt
cat gpurun_out/r2s5_conv_trace.txt
timeout 600 python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2s5_bench.json 2> gpurun_out/r2s5_bench.err
tail -c 1000 gpurun_out/r2s5_bench.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/r2s5_bench.json"))
print({k: j[k] for k in ("value", "ms_per_step", "stages_ms")}, j["e2e"]["value"], j["serial"])
PY
