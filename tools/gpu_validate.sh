#!/bin/bash
# Full GPU validation of a build: all -m gpu tests, smoke, compute-sanitizer over the hot kernels (small inputs), bench.
# usage (from the repo root):  gpurun --timeout 2400 -- 'bash tools/gpu_validate.sh [tests|sanitize|bench ...]'
mkdir -p gpurun_out
what="${@:-tests sanitize bench}"
for w in $what; do
  case $w in
    tests)
      timeout 900 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -15 > gpurun_out/validate_tests.log
      tail -4 gpurun_out/validate_tests.log
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ;;
    sanitize)
      K1='test_conv_matches_torch and (44 or 28 or 45) and not simt and not 60 and not 120'
      K2='(test_select_seeds_bit_exact and 32-48-64-100) or (test_full_clustering_matches_reference_golden and cluster_d) or test_assign_tensor_core_certificate_on_marginless_input or test_select_seeds_with_init_seeds or (test_hill_climb_vs_double_oracle and 32-48-64-100)'
      for tool in memcheck racecheck; do
        timeout 1200 compute-sanitizer --tool $tool --error-exitcode 86 python -m pytest tests/test_gpu_backbone.py -m gpu -q -x -k "$K1" > gpurun_out/sanitizer_${tool}_conv.log 2>&1
        echo "$tool conv rc=$?"; tail -4 gpurun_out/sanitizer_${tool}_conv.log
        timeout 1200 compute-sanitizer --tool $tool --error-exitcode 86 python -m pytest tests/test_gpu_clustering.py -m gpu -q -x -k "$K2" > gpurun_out/sanitizer_${tool}_cluster.log 2>&1
        echo "$tool cluster rc=$?"; tail -4 gpurun_out/sanitizer_${tool}_cluster.log
      done ;;
    bench)
      timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err
      tail -c 500 gpurun_out/validate_bench.err; head -c 1200 gpurun_out/validate_bench.json; echo ;;
  esac
done
