#!/bin/bash
# session-2 call D: loop trace with the issuing thread's waits; PDL A/B on the backbone; backbone tests with PDL
mkdir -p gpurun_out
for poly in 0 8; do
  UOC_LOOP_POLY=$poly timeout 120 python tools/loop_trace.py > /dev/null 2> gpurun_out/loop_trace_poly$poly.txt; echo "trace poly=$poly exit $?"; grep "update [45]" gpurun_out/loop_trace_poly$poly.txt | grep -v "P-1" | tail -4
done
for pdl in 0 1; do
  UOC_CONV_PDL=$pdl UOC_AB_TAG=_pdl$pdl timeout 300 python tools/batch_ab.py 1 2 > gpurun_out/batch_ab_pdl$pdl.log 2>&1; echo "pdl=$pdl exit $?"; tail -2 gpurun_out/batch_ab_pdl$pdl.log
done
timeout 900 python -m pytest tests/test_gpu_backbone.py -m gpu -q --timeout 300 -p no:cacheprovider -k "not conv_matches_torch or tc_cluster1 or tc_cluster4" > gpurun_out/t_backbone.log 2>&1; echo "backbone exit $?"; tail -3 gpurun_out/t_backbone.log
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_pipeline.log 2>&1; echo "pipeline exit $?"; tail -3 gpurun_out/t_pipeline.log
