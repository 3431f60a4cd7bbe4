#!/bin/bash
# compute-sanitizer over the persistent ring decode kernel (hand-rolled grid barrier, mbarrier ring, named barriers) and
# the encoder / frontend kernels of one micro clip.  Run on the GPU box:
#     gpurun --timeout 1500 -- 'bash tests/run_sanitizer.sh'
# Logs land in gpurun_out/sanitizer_*.log; the summaries are committed under profiles/.
#   memcheck  : out-of-bounds / misaligned global, shared and local accesses (incl. the bulk-copy destinations)
#   synccheck : divergent / mismatched bar.sync, bar.arrive and mbarrier use
#   racecheck : shared-memory hazards between the compute warps, the producer warp and the async proxy
# (racecheck does not model global memory: the release/relaxed protocol of the grid barrier is covered by the
#  determinism and N-rank equality tests and by memcheck's view of the data it orders.)
set -u
mkdir -p gpurun_out
export WM_SAN_ITERS=${WM_SAN_ITERS:-4}
rc=0
for tool in ${WM_SAN_TOOLS:-memcheck synccheck racecheck}; do
  log=gpurun_out/sanitizer_${tool}.log
  echo "== compute-sanitizer --tool $tool" | tee $log
  timeout ${WM_SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 77 \
      python tests/gpu_sanitize_target.py >> $log 2>&1
  r=$?
  echo "== exit code $r" | tee -a $log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|tokens|== exit" $log | tail -8
  if [ $r -ne 0 ] && [ $tool != racecheck ]; then rc=$r; fi   # racecheck: see profiles/r2_sanitizer.md (mbarrier-ordered ring slots)
done
exit $rc
