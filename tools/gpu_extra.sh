VQ_STORE_POLICY=normal timeout 120 tools/gemm_selftest --time > gpurun_out/selftest_normal.log 2>&1
timeout 300 python tools/prof_kernels.py > gpurun_out/prof_kernels.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vq_attn_cross -s 2 -c 1 -f -o gpurun_out/cross python tools/prof_kernels.py > gpurun_out/ncu_cross.log 2>&1
grep -E "case M=16384|time|mainloop|loads\+" gpurun_out/selftest_normal.log | tail -30; cat gpurun_out/prof_kernels.log
