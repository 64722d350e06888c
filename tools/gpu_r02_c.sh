#!/bin/bash
# Round-2 perf A/B: temporal attention + fused quantiser, reverse-order quantise passes, GEMM store policy.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_stdit.py -q -x > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_c.log
tail -4 gpurun_out/pytest_c.log
b() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-peak > gpurun_out/bench_c_$tag.log 2>&1
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
ls = [l for l in open(f"gpurun_out/bench_c_{tag}.log").read().splitlines() if l.startswith("{")]
if not ls:
    print(tag, "NO LINE"); raise SystemExit
l = json.loads(ls[-1])
cls = {r["kernel_class"]: round(r["ms_per_step_single_profiled_replay"], 2) for r in l["roofline_kernels"]}
print(f"{tag:28s} {l['ms_per_step']:.2f} ms/step  busy {l['step_busy_ms_cupti']:.2f}  {cls}  {l['clocks']['sm_mhz']} MHz")
PY
}
b base            VQ_AQ_REVERSE=0 VQ_TEMPORAL_FUSED_QUANT=0
b tq              VQ_AQ_REVERSE=0 VQ_TEMPORAL_FUSED_QUANT=1
b rev             VQ_AQ_REVERSE=1 VQ_TEMPORAL_FUSED_QUANT=0
b rev_normal      VQ_AQ_REVERSE=1 VQ_TEMPORAL_FUSED_QUANT=0 VQ_STORE_POLICY=normal
b all             VQ_AQ_REVERSE=1 VQ_TEMPORAL_FUSED_QUANT=1
b all_normal      VQ_AQ_REVERSE=1 VQ_TEMPORAL_FUSED_QUANT=1 VQ_STORE_POLICY=normal
