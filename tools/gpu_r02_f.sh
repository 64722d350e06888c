#!/bin/bash
# Residual-prefetch GEMM epilogue: correctness (self-test incl. ragged / in-place cases), timing A/B, step A/B.
mkdir -p gpurun_out
timeout 200 tools/gemm_selftest > gpurun_out/selftest.log 2>&1; echo "selftest rc=$?" >> gpurun_out/selftest.log
grep -E "epi=2|PASSED|FAILED|rc=" gpurun_out/selftest.log
for pf in 0 1; do
  VQ_GEMM_RESPF=$pf timeout 200 tools/gemm_selftest --time32 > gpurun_out/selftest32_respf$pf.log 2>&1
  echo "== respf=$pf"; grep -E "case|time " gpurun_out/selftest32_respf$pf.log | grep -A1 "epi=2"
done
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_stdit.py -q -x 2>&1 | tail -2
for pf in 0 1; do
  VQ_GEMM_RESPF=$pf timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-peak > gpurun_out/bench_respf$pf.log 2>&1
  echo "== respf=$pf"; grep '^{' gpurun_out/bench_respf$pf.log | tail -1 | cut -c1-230
done
