#!/bin/bash
# session-2 call Z: evaluation tail (metrics.cu): parity + timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout 300 -p no:cacheprovider -k "multilabel" > gpurun_out/t_metrics.log 2>&1; echo "metrics tests exit $?"; tail -15 gpurun_out/t_metrics.log
python - <<'PY'
import sys, time
sys.path.insert(0, 'oracle'); sys.path.insert(0, '.')
import numpy as np, torch
import uoc_oracle as O
from unseenobjectclustering_b200 import evaluation as EV
_, gt = O.synthetic_clustered_features(480, 640, 8, 6, 0.05, 204)
gt = gt.numpy().astype(np.float32)
pred = O.synthetic_prediction(gt.astype(np.int64), 4).astype(np.float32)
for _ in range(3): EV.multilabel_metrics(pred, gt)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10): EV.multilabel_metrics(pred, gt)
torch.cuda.synchronize(); t_gpu = (time.perf_counter() - t) / 10
t = time.perf_counter(); O.multilabel_metrics(pred, gt); t_cpu = time.perf_counter() - t
print("multilabel_metrics 640x480, 6 gt / 8 predicted objects: device path %.2f ms, reference arithmetic on the host %.1f ms" % (t_gpu * 1e3, t_cpu * 1e3))
PY
