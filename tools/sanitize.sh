#!/bin/bash
# compute-sanitizer passes over the small-geometry GPU parity tests (union-find CCL, deblend walk, FOCAS, TMA/mbarrier warp
# kernel, marching sky kernel).  Logs go to gpurun_out/; the summaries are copied into profiles/.
OUT=${OUT:-gpurun_out}
SEL_MEM='test_star_lists_identical or test_star_lists_edge_cases or test_alignment_matches_oracle or (test_warp_bit_exact and 332) or (test_warp_accumulate_many and 1) or (test_sky_subtraction and (203 or 17 or 130)) or test_stack_mono_darks_only or test_masters_and_calibration'
SEL_RACE='(test_star_lists_identical and 2) or (test_warp_bit_exact and 332 and 1) or (test_sky_subtraction and 203) or test_stack_mono_darks_only'
for tool in memcheck racecheck synccheck; do
  sel="$SEL_MEM"; [ $tool = racecheck ] && sel="$SEL_RACE"; [ $tool = synccheck ] && sel="$SEL_RACE"
  echo "== compute-sanitizer --tool $tool  (pytest -k \"$sel\")"
  timeout ${LIMIT:-900} compute-sanitizer --tool $tool --target-processes all --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$sel" > $OUT/sanitize_$tool.log 2>&1
  echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|========= (Invalid|Race|Barrier|Error)" $OUT/sanitize_$tool.log | sort | uniq -c | head -20
done
