#!/bin/bash
# session-2 call I: seed selection kernel experiments (tile skip: reverted; shared exact-distance rounds)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_clustering.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "select_seeds or golden or cluster or two" > gpurun_out/t_fps.log 2>&1; echo "fps tests exit $?"; tail -5 gpurun_out/t_fps.log
timeout 200 python tools/fps_tc_trace.py 0 70 2> gpurun_out/fps_tc_trace_share.txt; grep -E "pass (1|2|3|10|30|50|70|90):|mean" gpurun_out/fps_tc_trace_skip.txt | tail -20
timeout 300 python tools/batch_ab.py 1 > gpurun_out/batch_ab_share.log 2>&1; tail -1 gpurun_out/batch_ab_skip.log
