#!/bin/bash
# session-2 call X: screening margin for round-to-nearest bf16 copies (A/B)
mkdir -p gpurun_out
UOC_FPS_RN_MARGIN=1 timeout 900 python -m pytest tests/test_gpu_clustering.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "select_seeds or golden or config5" > gpurun_out/t_rn.log 2>&1; echo "rn tests exit $?"; tail -3 gpurun_out/t_rn.log
for v in 0 1; do
  UOC_FPS_RN_MARGIN=$v UOC_AB_TAG=_rn$v timeout 300 python tools/batch_ab.py 1 > gpurun_out/batch_ab_rn$v.log 2>&1; tail -1 gpurun_out/batch_ab_rn$v.log
done
UOC_FPS_RN_MARGIN=1 timeout 300 python tools/bench_configs.py 2>&1 | tail -1 | cut -c90-200
