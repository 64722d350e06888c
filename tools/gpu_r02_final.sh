#!/bin/bash
# Final round-2 record: GPU suite, smoke, bench lines of every workload / schedule, ncu launch list + traffic of the bench command.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu --format=csv > gpurun_out/smi.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=" gpurun_out/pytest_gpu.log | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
grep '^{' gpurun_out/bench.log | tail -1 | cut -c1-300
for wl in linear pixart512 w4a8mp w8a8static; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.log 2>&1
  echo "== $wl"; grep '^{' gpurun_out/bench_$wl.log | tail -1 | cut -c1-200
done
for wl in sample100 sample20mp; do
  timeout 600 python bench.py --workload $wl --steps 2 --warmup 1 > gpurun_out/bench_$wl.log 2>&1
  echo "== $wl"; grep '^{' gpurun_out/bench_$wl.log | tail -1 | cut -c1-200
done
timeout 600 python bench.py --attn-int8 --steps 5 --warmup 3 --no-cpu-baseline --no-peak > gpurun_out/bench_attn_int8.log 2>&1
echo "== attn-int8"; grep '^{' gpurun_out/bench_attn_int8.log | tail -1 | cut -c1-200
timeout 900 python bench.py --schedule hook-graph --steps 5 --warmup 3 --no-cpu-baseline --no-peak > gpurun_out/bench_hook_graph.log 2>&1
echo "== hook-graph"; grep '^{' gpurun_out/bench_hook_graph.log | tail -1 | cut -c1-200
VQ_PDL=1 timeout 600 python bench.py --workload pixart512 --steps 10 --warmup 3 --no-cpu-baseline --no-peak > gpurun_out/bench_pixart512_pdl.log 2>&1
echo "== pixart512 PDL"; grep '^{' gpurun_out/bench_pixart512_pdl.log | tail -1 | cut -c1-200
timeout 900 python bench.py --schedule hook --steps 5 --warmup 3 --no-cpu-baseline --no-peak > gpurun_out/bench_hook.log 2>&1
echo "== hook"; grep '^{' gpurun_out/bench_hook.log | tail -1 | cut -c1-200
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log | cut -c1-200
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 6000 --csv \
   --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 1 --no-graph --depth 2 --no-cpu-baseline --no-peak > gpurun_out/bench_ncu.log 2>&1
python tools/ncu_traffic.py gpurun_out/traffic.csv > gpurun_out/traffic_table.md 2>&1; head -12 gpurun_out/traffic_table.md
timeout 300 python tools/timeline.py --depth 4 > gpurun_out/timeline.log 2>&1; head -14 gpurun_out/timeline.log | cut -c1-160
rm -f gpurun_out/traffic.csv
