#!/bin/bash
# compute-sanitizer memcheck over the kernels added in session 2 (small cases): two-stage plumbing, evaluation counts,
# streaming sampler, euclidean stages, network variants (6-channel stem, head modes), CTA-pair convolution
mkdir -p gpurun_out
export PYTEST_ADDOPTS="-p no:cacheprovider"
run() {
  name=$1; shift
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file gpurun_out/sanitizer_$name.log "$@" > gpurun_out/sanitizer_$name.out 2>&1
  rc=$?
  echo "$name: exit $rc; $(grep -c 'Invalid\|out of bounds\|misaligned' gpurun_out/sanitizer_$name.log 2>/dev/null) suspicious lines; $(grep 'ERROR SUMMARY' gpurun_out/sanitizer_$name.log | tail -1)"; tail -1 gpurun_out/sanitizer_$name.out
}
run refine python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k "kernel_matches_oracle and not 480"
run metrics python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k "multilabel_counts"
run euclid python -m pytest tests/test_gpu_clustering.py -m gpu -q -x -k "euclidean_stages and 33"
run stream python -m pytest tests/test_gpu_clustering.py -m gpu -q -x -k "select_seeds_bit_exact and stream and 24-24"
run variants python -m pytest tests/test_gpu_backbone.py -m gpu -q -x -k "variant and early and tcgen05 or variant and cat and tcgen05"
run pair python -m pytest tests/test_gpu_backbone.py -m gpu -q -x -k "conv_matches_torch and pair and 20-28-2"
