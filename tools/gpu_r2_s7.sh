#!/bin/bash
# round 2, session 7: direct finishing + per-layer dispatch: parity, per-layer A/B, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backbone.py -m gpu -x -q --timeout 60 2>&1 | tail -5
UOC_CONV_LAYERS_VARIANTS="auto;UOC_CONV_PAIR=1;UOC_CONV_PAIR=0" timeout 300 python tools/conv_layers.py 2 8 > gpurun_out/r2s10_conv_layers.txt 2>&1
cat gpurun_out/r2s10_conv_layers.txt
UOC_CONV_LAYERS_ONLY="l3 3x3 256 d2,l4 3x3 512 d4" UOC_CONV_LAYERS_VARIANTS="UOC_CONV_TRACE=1,UOC_CONV_PAIR=1" timeout 300 python tools/conv_layers.py 2 8 2>&1 | grep "conv trace" | awk 'NR%7==1' > gpurun_out/r2s10_conv_trace.txt
cat gpurun_out/r2s10_conv_trace.txt
timeout 600 python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2s10_bench.json 2> gpurun_out/r2s10_bench.err
tail -c 1000 gpurun_out/r2s10_bench.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/r2s10_bench.json"))
print({k: j[k] for k in ("value", "ms_per_step", "stages_ms")}, j["e2e"]["value"], j["serial"])
PY
