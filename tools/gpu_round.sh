#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 300 tools/gemm_selftest --time > gpurun_out/selftest.log 2>&1; echo "selftest rc=$?" >> gpurun_out/selftest.log
VQ_STORE_POLICY=normal timeout 300 tools/gemm_selftest --time > gpurun_out/selftest_normal.log 2>&1
timeout 300 python tools/prof_kernels.py > gpurun_out/prof_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/prof_kernels.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
VQ_STORE_POLICY=normal timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_normal.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
echo "== evict-first stores"; grep -E "case M=16384|time|mainloop|loads|PASSED|FAILED|mismatches [1-9]" gpurun_out/selftest.log
echo "== normal stores"; grep -E "time |loads\+stores" gpurun_out/selftest_normal.log
grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu.log | tail -5; cat gpurun_out/prof_kernels.log; tail -2 gpurun_out/bench.log | cut -c1-400; tail -2 gpurun_out/bench_normal.log | cut -c1-300
