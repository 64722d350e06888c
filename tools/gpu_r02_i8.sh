#!/bin/bash
# INT8 attention: ncu launch metrics of the three kernels + one --set full capture of the attention kernel
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
cat > /tmp/i8_once.py <<'P'
import sys, torch
sys.path.insert(0, ".")
from viditq_b200 import ops
n_seq, S, H, D = 32, 1024, 16, 72
x = torch.randn(n_seq * S, 3 * H * D, device="cuda").half()
for _ in range(2):
    o = ops.attn_spatial_i8(x, n_seq, S, H, D, D ** -0.5)
torch.cuda.synchronize()
P
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,smsp__inst_executed.sum --clock-control none -k regex:"ia_|attn_i8" --csv --log-file gpurun_out/i8_launches.csv python /tmp/i8_once.py > gpurun_out/i8_ncu1.log 2>&1
grep -E "ia_|attn_i8" gpurun_out/i8_launches.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | tail -20
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vq_attn_i8 -s 1 -c 1 -f -o gpurun_out/attn_i8 python /tmp/i8_once.py > gpurun_out/i8_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
