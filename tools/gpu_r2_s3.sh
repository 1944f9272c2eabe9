#!/bin/bash
# round 2, session 3: the persistent CTA-pair stream-K convolution + the row-staged head: parity, per-layer timings, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backbone.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2s3_tests.log
tail -15 gpurun_out/r2s3_tests.log
UOC_CONV_LAYERS_VARIANTS="pair;UOC_CONV_PAIR=0" timeout 300 python tools/conv_layers.py 2 8 > gpurun_out/r2s3_conv_layers.txt 2>&1
cat gpurun_out/r2s3_conv_layers.txt
timeout 600 python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2s3_bench.json 2> gpurun_out/r2s3_bench.err
tail -c 1500 gpurun_out/r2s3_bench.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/r2s3_bench.json"))
print({k: j[k] for k in ("value", "ms_per_step", "stages_ms")}, j["e2e"]["value"], j["serial"])
PY
