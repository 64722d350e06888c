#!/bin/bash
# Multi-GPU round: 2-GPU bit-identity tests (eager + graph segments), strong-scaling bench lines.  $1 = number of GPUs leased.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -5 gpurun_out/pytest_multi.log
run() {  # n parallelism extra...
  n=$1; par=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
     bench.py --gpus $n --steps 8 --warmup 3 --no-cpu-baseline --no-peak --parallelism $par "$@" > gpurun_out/bench_n${n}_${par}$TAG.log 2>&1
  echo "== n=$n $par $TAG rc=$?"; grep '^{' gpurun_out/bench_n${n}_${par}$TAG.log | tail -1 | cut -c1-330
  grep -o '"multi_gpu_parity": {[^}]*}' gpurun_out/bench_n${n}_${par}$TAG.log | tail -1
}
TAG=""
run 2 cfg-branch
run 2 frames
TAG="_eager" run 2 frames --no-graph
if [ "$N" -ge 4 ]; then run 4 frames; fi
if [ "$N" -ge 8 ]; then run 8 frames; run 8 cfg-branch; fi
run $N samples
