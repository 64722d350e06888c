timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --cfg-mode split > gpurun_out/bench_split.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --cfg-mode stacked > gpurun_out/bench_stacked2.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --cfg-mode split > gpurun_out/bench_split2.log 2>&1
for f in bench_split bench_stacked2 bench_split2; do tail -1 gpurun_out/$f.log | cut -c1-200; done
