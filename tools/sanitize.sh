#!/bin/bash
# compute-sanitizer over the hand-rolled mbarrier / TMEM / TMA kernels (SURVEY.md §5 aux: race + memory checking).
# memcheck on the GEMM / attention self-tests and the fused-linear parity cases; racecheck (shared-memory hazards between
# the producer / MMA / epilogue warps) on the small cases only — it serialises every shared access.
# Usage (GPU box):  bash tools/sanitize.sh  > gpurun_out/sanitize.log
set -u
cd "$(dirname "$0")/.."
CS=${COMPUTE_SANITIZER:-compute-sanitizer}
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | tail -15; echo "rc=${PIPESTATUS[0]}"; }
run $CS --tool memcheck --error-exitcode 9 tools/gemm_selftest
run $CS --tool memcheck --error-exitcode 9 tools/attn_selftest
run $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_linear.py -q -x -k "109 or 300 or 1024-1152 or oracle"
run $CS --tool racecheck --racecheck-report hazard --error-exitcode 9 tools/gemm_selftest --case 256 384 1152 2
run $CS --tool racecheck --racecheck-report hazard --error-exitcode 9 python -m pytest tests/test_gpu_linear.py -q -x -k "109 or 300"
run $CS --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_linear.py -q -x -k "109 or 2-1024-1152"
