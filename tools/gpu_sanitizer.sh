#!/bin/bash
# compute-sanitizer over the embedding path's kernels (K1 gather / pool, K2 sort + fused update, C1 pack / scatter) and
# the K6 fused GEMMs: memcheck (out-of-bounds / misaligned accesses) and racecheck (shared-memory hazards) on small
# cases of the GPU suite.  Output: gpurun_out/r2_sanitizer_{memcheck,racecheck}.log (summaries copied to profiles/).
# NOTE (round 2): this GPU pool refuses compute-sanitizer ("closed on this pool ...", exit code 86; the refusal is kept in
# profiles/r2_sanitizer_refused.txt), so the script is for a box that allows it; on this pool the substitutes are the
# kernels' own checks (bounded mbarrier waits that trap, the out-of-range id flag, list-overflow and scale-overflow
# words) and the property tests of tests/test_properties.py.
mkdir -p gpurun_out
SEL='test_onehot_gather_is_bit_exact or test_bag_pooling_matches_reference_idioms or test_sort_dedup_bit_exact_onehot or test_fused_update or test_pack or test_scatter or test_index_prep_bit_exact'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
      python -m pytest tests/test_gpu_kernels.py -x -q -k "$SEL" -p no:cacheprovider \
      > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/r2_sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2_sanitizer_$tool.log | tail -n 4
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
    python -m pytest tests/test_gpu_tc_fused.py -x -q -p no:cacheprovider > gpurun_out/r2_sanitizer_memcheck_k6.log 2>&1
echo "memcheck k6 rc=$?" | tee -a gpurun_out/r2_sanitizer_memcheck_k6.log
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_sanitizer_memcheck_k6.log | tail -n 3
