#!/bin/bash
# round 2, session 6: seed selection of a BATCH of fields: taking turns in the resident-slice kernel vs streaming side by side
mkdir -p gpurun_out
for mode in 0 1; do
  echo "UOC_FPS_BATCH_STREAM=$mode"
  UOC_FPS_BATCH_STREAM=$mode UOC_AB_TAG=_fpsbatch$mode timeout 300 python tools/batch_ab.py 1 2 4 2>&1 | tail -4
done > gpurun_out/r2s6_fps_batch.txt 2>&1
cat gpurun_out/r2s6_fps_batch.txt
timeout 300 python -m pytest tests/test_gpu_clustering.py -m gpu -x -q --timeout 120 -k "select_seeds or batched or full_size or stale" 2>&1 | tail -5
