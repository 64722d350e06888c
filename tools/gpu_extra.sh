timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
rm -f gpurun_out/*.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --depth 2 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
for c in "16384 3456 1152 0" "16384 1152 1152 2" "16384 4608 1152 0" "16384 1152 4608 2"; do
  n=$(echo $c | tr ' ' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:vq_gemm -c 1 -f -o gpurun_out/gemm_$n \
     tools/gemm_selftest --case $c > gpurun_out/ncu_gemm_$n.log 2>&1
done
tail -2 gpurun_out/smoke.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vq_attn_spatial --launch-skip 30 -c 1 -f -o gpurun_out/attn_spatial \
   tools/attn_selftest --modes 0 --time > gpurun_out/ncu_attn.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:act_quant --launch-skip 3 -c 3 -f -o gpurun_out/quant \
   python tools/prof_quant.py > gpurun_out/ncu_quant.log 2>&1
timeout 300 python tools/prof_kernels.py > gpurun_out/prof_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/prof_kernels.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu --format=csv > gpurun_out/smi.txt
tail -30 gpurun_out/prof_kernels.log
