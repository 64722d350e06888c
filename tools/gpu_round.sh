#!/bin/bash
# Correctness + timing round: selftest, GPU parity tests, in-graph timeline, bench.
mkdir -p gpurun_out
timeout 120 tools/gemm_selftest --time > gpurun_out/selftest.log 2>&1; echo "selftest rc=$?" >> gpurun_out/selftest.log
timeout 120 tools/attn_selftest --time > gpurun_out/attn_selftest.log 2>&1; echo "attn_selftest rc=$?" >> gpurun_out/attn_selftest.log
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/timeline.py --depth 4 > gpurun_out/timeline.log 2>&1; echo "rc=$?" >> gpurun_out/timeline.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
if [ -n "$VQ_ROUND_EXTRA" ]; then bash -c "$VQ_ROUND_EXTRA"; fi
grep -E "case M=16384|time|mainloop|loads\+|PASSED|FAILED|rc=" gpurun_out/selftest.log | tail -30
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -5; head -24 gpurun_out/timeline.log | cut -c1-150; tail -2 gpurun_out/bench.log | cut -c1-400
