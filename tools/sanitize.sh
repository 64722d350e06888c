#!/bin/bash
# compute-sanitizer over the hand-rolled mbarrier / TMEM / TMA kernels (SURVEY.md §5 aux: race + memory checking).
# memcheck and synccheck on the GEMM / attention self-tests and the fused-linear parity cases; racecheck (shared-memory
# hazards between the producer / MMA / converter / epilogue warps) on small cases only — it serialises every shared access.
# racecheck does not model mbarrier arrive (release) / try_wait (acquire) hand-offs, so every mbarrier-synchronised
# buffer shows up as a "potential hazard": the script prints the distinct (writer, reader) source-line pairs so that each
# can be matched to its barrier (profiles/r02_s6_sanitize.md).
# Usage (GPU box):  bash tools/sanitize.sh  > gpurun_out/sanitize.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=${COMPUTE_SANITIZER:-compute-sanitizer}
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|SELFTEST|mismatch" | tail -6; echo "rc=${PIPESTATUS[0]}"; }
race() {
  tag=$1; shift
  echo "=== racecheck $*"
  timeout 900 $CS --tool racecheck --racecheck-report hazard --print-limit 100000 "$@" > /tmp/racecheck_$tag.log 2>&1
  echo "rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|mismatches" /tmp/racecheck_$tag.log | tail -3
  # distinct hazard kinds: "<type> | writer file:line | reader file:line"
  awk '/hazard detected/ {t=$4} /Write Thread/ {w=$NF} /Read Thread/ {print t " | write " w " | read " $NF}' /tmp/racecheck_$tag.log \
    | sort | uniq -c | sort -rn | head -30
}
run $CS --tool memcheck --error-exitcode 9 tools/gemm_selftest
run $CS --tool memcheck --error-exitcode 9 tools/attn_selftest
run $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_linear.py -q -x -k "109 or 300 or 1024-1152 or oracle or packed"
run $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "temporal_attention_with_fused"
run $CS --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_linear.py -q -x -k "109 or 2-1024-1152"
race gemm tools/gemm_selftest --case 256 384 1152 2
race linear python -m pytest tests/test_gpu_linear.py -q -x -k "1-109-2304 or 1-300-1152-2"
# opt-in INT8 attention (vq_attn_i8.cu): operand passes + the two-pass tcgen05 kind::i8 kernel, small shapes
run $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_attn_i8.py -q -x -m gpu -k "256-1 or 512-2 or refuses"
run $CS --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_attn_i8.py -q -x -m gpu -k "256-1"
