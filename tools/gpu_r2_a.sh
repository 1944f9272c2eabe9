#!/bin/bash
# session-2 call A: network variants, loop A/B (MUFU / FMA split), clustering tests under each split
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
timeout 600 python -m pytest tests/test_gpu_backbone.py -m gpu -q --timeout 300 -p no:cacheprovider -k "variant or golden or drop_in or full_frame" > gpurun_out/t_variants.log 2>&1; echo "variants exit $?"; tail -15 gpurun_out/t_variants.log
timeout 300 python tools/loop_ab.py > gpurun_out/loop_ab.log 2>&1; echo "loop_ab exit $?"; tail -14 gpurun_out/loop_ab.log
for poly in 8 12; do
  UOC_LOOP_POLY=$poly timeout 600 python -m pytest tests/test_gpu_clustering.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/t_cluster_poly$poly.log 2>&1; echo "clustering poly=$poly exit $?"; tail -3 gpurun_out/t_cluster_poly$poly.log
done
