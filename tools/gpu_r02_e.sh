#!/bin/bash
# Validation round: full GPU suite, smoke, sanitizer, default bench, rasterisation group A/B, per-kernel timings.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1; grep -E "^===|ERROR SUMMARY|RACECHECK SUMMARY|rc=| write " gpurun_out/sanitize.log | head -60
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
grep '^{' gpurun_out/bench.log | tail -1 | cut -c1-700
for gm in 4 16; do
  VQ_GEMM_GROUP_M=$gm timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-peak > gpurun_out/bench_gm$gm.log 2>&1
  echo "== group_m=$gm"; grep '^{' gpurun_out/bench_gm$gm.log | tail -1 | cut -c1-230
done
timeout 300 python tools/prof_kernels.py > gpurun_out/prof_kernels.log 2>&1; grep -E "patch_embed|act_quant|attn_temporal" gpurun_out/prof_kernels.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log | cut -c1-400
