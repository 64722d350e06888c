#!/bin/bash
# Round-2 second GPU call: W4 packed path + chunked MLP correctness, workload / schedule bench lines, MLP-chunk A/B,
# ncu traffic at the launched shapes, ncu --set full captures of the GEMM shapes at M = 32768.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linear.py tests/test_gpu_deep.py -q -x -s > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_b.log
grep -E "passed|failed|FAILED|Error|rc=|DDIM|simulation on|schedule" gpurun_out/pytest_b.log | tail -20
timeout 300 python -m pytest tests/test_gpu_stdit.py -q -x -k "chunked or noise_band" > gpurun_out/pytest_b2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_b2.log
tail -3 gpurun_out/pytest_b2.log
for wl in linear pixart512 w4a8mp; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.log 2>&1; echo "rc=$?" >> gpurun_out/bench_$wl.log
  tail -2 gpurun_out/bench_$wl.log | cut -c1-1200
done
VQ_LINEAR_FUSED=0 timeout 600 python bench.py --workload pixart512 --steps 10 --warmup 3 --no-cpu-baseline --no-peak > gpurun_out/bench_pixart512_twolaunch.log 2>&1
tail -1 gpurun_out/bench_pixart512_twolaunch.log | cut -c1-400
VQ_LINEAR_FUSED=1 timeout 600 python bench.py --workload linear --steps 10 --warmup 3 --no-cpu-baseline --no-peak > gpurun_out/bench_linear_fused16k.log 2>&1
tail -1 gpurun_out/bench_linear_fused16k.log | cut -c1-400
timeout 900 python bench.py --schedule hook --steps 5 --warmup 3 --no-cpu-baseline --no-peak > gpurun_out/bench_hook.log 2>&1; echo "rc=$?" >> gpurun_out/bench_hook.log
tail -2 gpurun_out/bench_hook.log | cut -c1-600
for ch in 3072 4096 6144; do
  VQ_MLP_CHUNK=$ch timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-peak > gpurun_out/bench_chunk$ch.log 2>&1
  echo "== mlp chunk $ch"; tail -1 gpurun_out/bench_chunk$ch.log | cut -c1-260
done
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 6000 --csv \
   --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 1 --no-graph --depth 2 --no-cpu-baseline --no-peak > gpurun_out/bench_ncu.log 2>&1
python tools/ncu_traffic.py gpurun_out/traffic.csv > gpurun_out/traffic_table.md 2>&1; tail -25 gpurun_out/traffic_table.md
rm -f gpurun_out/*.ncu-rep
for c in "32768 1152 1152 2" "32768 3456 1152 0" "32768 1152 4608 2"; do
  n=$(echo $c | tr ' ' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:vq_gemm -c 1 -f -o gpurun_out/gemm_$n \
     tools/gemm_selftest --case $c > gpurun_out/ncu_gemm_$n.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
