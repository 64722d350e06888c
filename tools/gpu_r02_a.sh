#!/bin/bash
# Round-2 first GPU call: correctness of everything new, then numbers.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu --format=csv > gpurun_out/smi.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 200 tools/gemm_selftest > gpurun_out/selftest.log 2>&1; echo "selftest rc=$?" >> gpurun_out/selftest.log
tail -2 gpurun_out/selftest.log
timeout 600 python -m pytest tests/test_gpu_linear.py -q -x > gpurun_out/pytest_linear.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_linear.log
tail -5 gpurun_out/pytest_linear.log
timeout 1500 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_linear.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|rel-|band|parity|steps|rc=" gpurun_out/pytest_gpu.log | tail -60
for gm in 0 8; do
  VQ_GEMM_GROUP_M=$gm timeout 200 tools/gemm_selftest --time32 > gpurun_out/selftest32_gm$gm.log 2>&1
  echo "== group_m=$gm"; grep -E "case|time " gpurun_out/selftest32_gm$gm.log
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-3000
VQ_GEMM_GROUP_M=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-peak > gpurun_out/bench_gm0.log 2>&1
tail -1 gpurun_out/bench_gm0.log | cut -c1-300
