#!/bin/bash
# compute-sanitizer over a reduced GPU test subset (the hand-rolled mbarrier / TMEM
# protocols and the atomic append buffers are what can race or run out of bounds).
#   memcheck   out-of-bounds / misaligned global + shared accesses, leaks of device allocations
#   racecheck  shared-memory hazards (the decoders' generic-proxy stores vs the UMMA reads are ordered
#              by fence.proxy.async + mbarriers, which racecheck does not model: hazards it reports
#              on the swizzled operand stages are listed, not failed on)
#   synccheck  barrier misuse
# Summaries -> gpurun_out/r2_sanitizer_*.txt (copy into profiles/).  Run under gpurun, 1 GPU.
set -x
SUBSET="tests/test_gpu_flat.py::test_exact_search_parity tests/test_gpu_flat.py::test_merge_topk_known_answers_on_gpu \
tests/test_gpu_ivf.py::test_ivf_flat_recall_matches_oracle tests/test_gpu_ivf.py::test_ivf_pq_recall_matches_oracle \
tests/test_gpu_ivf.py::test_grouped_scan_seed_modes_agree_on_structureless_data \
tests/test_gpu_comm.py::test_single_rank_communicator_is_the_plain_search"
for tool in memcheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
      python -m pytest $SUBSET -m gpu -x -q > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" | tee gpurun_out/r2_sanitizer_$tool.rc
  grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|Hazard" gpurun_out/r2_sanitizer_$tool.log | tail -8 > gpurun_out/r2_sanitizer_$tool.txt
done
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 \
    python -m pytest tests/test_gpu_ivf.py::test_ivf_flat_recall_matches_oracle tests/test_gpu_flat.py::test_merge_topk_known_answers_on_gpu \
    -m gpu -x -q > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" | tee gpurun_out/r2_sanitizer_racecheck.rc
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2_sanitizer_racecheck.log | sort | uniq -c | sort -rn | head -12 > gpurun_out/r2_sanitizer_racecheck.txt
