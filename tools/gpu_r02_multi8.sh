#!/bin/bash
# 8-GPU strong-scaling record: one sample's denoise step frame-sharded over 2 / 4 / 8 GPUs, cfg-branch pairs at 2 and 8.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
run() {
  n=$1; par=$2; shift 2
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
     bench.py --gpus $n --steps 8 --warmup 3 --no-cpu-baseline --no-peak --parallelism $par "$@" > gpurun_out/bench_n${n}_${par}.log 2>&1
  echo "== n=$n $par rc=$?"; grep '^{' gpurun_out/bench_n${n}_${par}.log | tail -1 | cut -c1-200
  grep -o '"multi_gpu_parity": {[^}]*}' gpurun_out/bench_n${n}_${par}.log | tail -1
}
run 8 frames
run 4 frames
run 2 frames --no-selfcheck
run 2 cfg-branch --no-selfcheck
run 8 cfg-branch --no-selfcheck
