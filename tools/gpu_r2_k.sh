#!/bin/bash
# session-2 call K: CTA-pair convolution (cta_group::2): parity, then backbone timing with / without
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_backbone.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k "conv_matches_torch and pair" > gpurun_out/t_pair.log 2>&1; echo "pair tests exit $?"; tail -25 gpurun_out/t_pair.log
for v in 0 1; do
  UOC_CONV_2SM=$v UOC_AB_TAG=_2sm$v timeout 300 python tools/batch_ab.py 1 4 > gpurun_out/batch_ab_2sm$v.log 2>&1; echo "2sm=$v exit $?"; tail -2 gpurun_out/batch_ab_2sm$v.log
done
