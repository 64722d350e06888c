#!/bin/bash
# One batched GPU session. Outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 300 tools/gemm_selftest --time > gpurun_out/selftest.log 2>&1; echo "selftest rc=$?" >> gpurun_out/selftest.log
timeout 300 tools/gemm_selftest_e8 --time > gpurun_out/selftest_e8.log 2>&1; echo "selftest rc=$?" >> gpurun_out/selftest_e8.log
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/prof_kernels.py > gpurun_out/prof_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/prof_kernels.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_gemm_w8a8_kernel -s 5 -c 1 -f -o gpurun_out/gemm_v6_sq \
   python tools/prof_kernels.py > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_act_quant -s 2 -c 2 -f -o gpurun_out/actq3 \
   python tools/prof_kernels.py > gpurun_out/ncu_actq.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --depth 2 > gpurun_out/bench_ncu.log 2>&1
echo "=== 16 warps"; grep -E "time|mainloop|PASSED|FAILED|mismatches [1-9]" gpurun_out/selftest.log
echo "=== 8 warps"; grep -E "time|mainloop|PASSED|FAILED|mismatches [1-9]" gpurun_out/selftest_e8.log
grep -E "passed|failed|FAILED|sim fp16|int layerwise|W4A8 smooth|cross-back|pixart" gpurun_out/pytest_gpu.log | tail -20; cat gpurun_out/prof_kernels.log; tail -3 gpurun_out/bench.log
