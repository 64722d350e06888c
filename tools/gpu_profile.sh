#!/bin/bash
# Profiling round: in-graph timeline, per-kernel timings, ncu launch list of the bench command, full captures of the
# dominant GEMM shapes.  Outputs under gpurun_out/ (summarised into profiles/ by tools/summarize_ncu.py).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 300 python tools/timeline.py --depth 4 > gpurun_out/timeline.log 2>&1; echo "rc=$?" >> gpurun_out/timeline.log
timeout 300 python tools/prof_kernels.py > gpurun_out/prof_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/prof_kernels.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --depth 2 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
for c in "16384 3456 1152 0" "16384 1152 1152 2" "16384 4608 1152 1" "16384 1152 4608 2"; do
  n=$(echo $c | tr ' ' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:vq_gemm -c 1 -f -o gpurun_out/gemm_$n \
     tools/gemm_selftest --case $c > gpurun_out/ncu_gemm_$n.log 2>&1
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu --format=csv > gpurun_out/smi.txt
cat gpurun_out/timeline.log; cat gpurun_out/prof_kernels.log
