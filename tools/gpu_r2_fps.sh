#!/bin/bash
# session-2: compaction of the exact-distance evaluations in the seed selection kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clustering.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/t_cluster.log 2>&1; echo "cluster exit $?"; tail -3 gpurun_out/t_cluster.log
timeout 300 python tools/fps_tc_stats.py 2> gpurun_out/fps_tc_stats_compact.txt; grep -E "^---|passes 1" gpurun_out/fps_tc_stats_compact.txt
timeout 300 python tools/batch_ab.py 1 2>&1 | tail -1
timeout 300 python tools/two_stage_profile.py 2>&1 | grep -E "clustering|sum"
