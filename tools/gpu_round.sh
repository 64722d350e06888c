#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 120 tools/gemm_selftest --time > gpurun_out/selftest.log 2>&1; echo "selftest rc=$?" >> gpurun_out/selftest.log
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/prof_kernels.py > gpurun_out/prof_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/prof_kernels.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_act_quant -s 2 -c 1 -f -o gpurun_out/actq6 \
   python tools/prof_kernels.py > gpurun_out/ncu_actq.log 2>&1
grep -E "case M=16384|time|mainloop|loads\+|PASSED|FAILED|rc=" gpurun_out/selftest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -5; cat gpurun_out/prof_kernels.log; tail -2 gpurun_out/bench.log | cut -c1-330
