#!/bin/bash
# session-2 call C: euclidean metric tests, full clustering + backbone tests, loop trace with / without the polynomial split
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_clustering.py -m gpu -q --timeout 300 -p no:cacheprovider -k "euclid" > gpurun_out/t_euclid.log 2>&1; echo "euclid exit $?"; tail -25 gpurun_out/t_euclid.log
timeout 600 python -m pytest tests/test_gpu_clustering.py -m gpu -q --timeout 300 -p no:cacheprovider -k "not euclid" > gpurun_out/t_cluster.log 2>&1; echo "cluster exit $?"; tail -3 gpurun_out/t_cluster.log
for poly in 0 8; do
  UOC_LOOP_POLY=$poly timeout 120 python tools/loop_trace.py > /dev/null 2> gpurun_out/loop_trace_poly$poly.txt; echo "trace poly=$poly exit $?"; grep "cta 0 update [45]" gpurun_out/loop_trace_poly$poly.txt | tail -2
done
