#!/bin/bash
mkdir -p gpurun_out
UOC_FPS_PREFETCH=1 timeout 600 python -m pytest tests/test_gpu_clustering.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "select_seeds or golden" > gpurun_out/t_pf.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/t_pf.log
for v in 0 1; do
  UOC_FPS_PREFETCH=$v timeout 300 python tools/batch_ab.py 1 2>&1 | tail -1
  UOC_FPS_PREFETCH=$v timeout 300 python tools/two_stage_profile.py 2>&1 | grep -E "^clustering"
done
# then the full validation of the tree as it is (prefetch off by default)
bash tools/gpu_final_s2.sh
