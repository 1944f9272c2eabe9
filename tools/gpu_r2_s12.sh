#!/bin/bash
mkdir -p gpurun_out
for deep in 0 1; do
  echo "UOC_FPS_DEEP=$deep"
  UOC_FPS_DEEP=$deep timeout 300 python tools/fps_stats.py 2>&1 | grep "fps stats" | awk 'NR%2==0'
  UOC_FPS_DEEP=$deep timeout 300 python tools/batch_ab.py 1 2>&1 | tail -1
done > gpurun_out/r2s12_fps_deep.txt 2>&1
cat gpurun_out/r2s12_fps_deep.txt
timeout 600 python -m pytest tests/test_gpu_clustering.py -m gpu -x -q --timeout 120 2>&1 | tail -3
