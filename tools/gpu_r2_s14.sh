#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_clustering.py -m gpu -x -q --timeout 120 2>&1 | tail -3
timeout 300 python tools/fps_stats.py 2>&1 | grep "fps stats"
timeout 300 python tools/batch_ab.py 1 4 2>&1 | tail -2
