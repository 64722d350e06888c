mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -4 gpurun_out/pytest_multi.log; grep -E "^E " gpurun_out/pytest_multi.log | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --no-peak --parallelism frames > gpurun_out/bench_n2_frames.log 2>&1
grep '^{' gpurun_out/bench_n2_frames.log | tail -1 | cut -c1-330; grep -o '"multi_gpu_parity": {[^}]*}' gpurun_out/bench_n2_frames.log | tail -1
