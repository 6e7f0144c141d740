#!/bin/bash
# compute-sanitizer over small fused steps of every kernel family (race / memory checking).
#   tools/sanitize.sh memcheck|racecheck|synccheck|initcheck [wire|siren|wide|multiscale|all]
# ONE tool per GPU call (B200_PROFILING.md: several tools in one call have left GPUs unusable).  Output: gpurun_out/sanitize_<tool>.log
set -u
tool=${1:-memcheck}
what=${2:-all}
mkdir -p gpurun_out
log=gpurun_out/sanitize_${tool}_${what}.log
# plain run first: a faulting program must not be handed to the sanitizer
timeout 120 python tools/sanitize_target.py "$what" 2 > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
timeout 900 compute-sanitizer --tool "$tool" --print-limit 20 --launch-timeout 0 python tools/sanitize_target.py "$what" 2 > "$log" 2>&1
rc=$?
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|steps ok|========= (Invalid|Race|Barrier|Uninit)" "$log" | head -40
echo "exit $rc (log: $log)"
