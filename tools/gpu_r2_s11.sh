#!/bin/bash
# round 2, session 11: speculative seed chain in the sampling kernel: bit-exactness, timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_clustering.py -m gpu -x -q --timeout 120 2>&1 | tail -8
timeout 300 python tools/batch_ab.py 1 4 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2s11_bench.json 2> gpurun_out/r2s11_bench.err
tail -c 600 gpurun_out/r2s11_bench.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/r2s11_bench.json"))
print({k: j[k] for k in ("value", "ms_per_step", "stages_ms")}, j["e2e"]["value"], j["serial"], j["config2_clustered"])
PY
