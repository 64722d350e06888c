#!/bin/bash
# One batched GPU session. Outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 300 tools/gemm_selftest --time > gpurun_out/selftest.log 2>&1; echo "selftest rc=$?" >> gpurun_out/selftest.log
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_nograph.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --depth 2 > gpurun_out/bench_ncu.log 2>&1
tail -22 gpurun_out/selftest.log; tail -12 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -5 gpurun_out/bench.log; tail -3 gpurun_out/bench_nograph.log
