#!/bin/bash
# One batched GPU session: self-test + timing, pytest -m gpu, ncu captures. Outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 300 tools/gemm_selftest --time > gpurun_out/selftest.log 2>&1; echo "selftest rc=$?" >> gpurun_out/selftest.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for c in "16384 1152 1152 0" "16384 1152 4608 0" "16384 4608 1152 1"; do
  tag=$(echo $c | tr ' ' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:vq_gemm -c 1 -f -o gpurun_out/gemm_$tag \
     tools/gemm_selftest --case $c > gpurun_out/ncu_$tag.log 2>&1
done
tail -30 gpurun_out/selftest.log; tail -15 gpurun_out/pytest_gpu.log
