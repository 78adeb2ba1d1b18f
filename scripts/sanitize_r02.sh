#!/bin/sh
# compute-sanitizer over the small parity shapes (VERDICT r1 missing #4):
#   memcheck  -- out-of-bounds / misaligned accesses, leaks of device allocations
#   racecheck -- shared-memory hazards (the lattice kernels exchange tiles through shared memory under __syncwarp)
#   synccheck -- divergent barriers
# Run under gpurun from the repo root; logs land in gpurun_out/ and are summarised into profiles/ by hand.
set -x
SEL='test_pairs_match_oracle or test_fused_estep_matches_oracle or test_grouped_centres_one_call or test_underflow_falls_back_in_gpu or test_split_tiles_match_one_centre_per_tile or test_streamed_estep_is_bit_identical_to_resident'
for tool in memcheck racecheck synccheck; do
  compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit code: $?" >> gpurun_out/sanitizer_$tool.log
  tail -4 gpurun_out/sanitizer_$tool.log
done
compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 \
  python -m pytest tests/test_em_loop.py tests/test_fb.py -x -q -m gpu > gpurun_out/sanitizer_memcheck_em_fb.log 2>&1
echo "memcheck (EM loop, forward-backward) exit code: $?" >> gpurun_out/sanitizer_memcheck_em_fb.log
tail -4 gpurun_out/sanitizer_memcheck_em_fb.log
