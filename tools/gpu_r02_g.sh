#!/bin/bash
# Overlapped schedule (quantiser warpgroups inside the persistent GEMM): parity, per-shape timing, step A/B.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_linear.py -q -x -k "overlapped" > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_g.log
tail -5 gpurun_out/pytest_g.log
timeout 300 python tools/overlap_bench.py > gpurun_out/overlap_bench.md 2>&1; cat gpurun_out/overlap_bench.md | tail -14
for m in 0 2; do
  VQ_LINEAR_FUSED=$m timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-peak > gpurun_out/bench_overlap$m.log 2>&1
  echo "== VQ_LINEAR_FUSED=$m"; grep '^{' gpurun_out/bench_overlap$m.log | tail -1 | cut -c1-230; tail -3 gpurun_out/bench_overlap$m.log | grep -v '^{' | cut -c1-200
done
VQ_LINEAR_FUSED=2 timeout 300 python -m pytest tests/test_gpu_stdit.py tests/test_gpu_deep.py -q -x 2>&1 | tail -3
