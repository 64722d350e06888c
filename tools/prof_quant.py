"""Driver for ncu captures of the HBM-bound quantise passes at the bench shapes (stacked cfg_split step, M = 32768).

    ncu --set full --clock-control none --import-source on -k regex:act_quant -c 6 -f -o gpurun_out/quant python tools/prof_quant.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from viditq_b200 import ops  # noqa: E402

M, C = 32768, 1152
torch.manual_seed(0)
x = (torch.randn(1, M, C, device="cuda") * 1.5).half()
h = (torch.randn(1, M, 4 * C, device="cuda") * 1.5).half()
shift = (torch.randn(2, C, device="cuda") * 0.1).half()
scale = (torch.randn(2, C, device="cuda") * 0.1).half()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(2):
    flush.zero_()
    ops.act_quant(x)                                                   # plain K = 1152
    flush.zero_()
    ops.ln_modulate_act_quant(x, shift, scale, rows_per_mod=M // 2)    # LayerNorm + modulate + quantise
    flush.zero_()
    ops.act_quant(h, gelu=True)                                        # GELU + quantise, K = 4608
torch.cuda.synchronize()
print("ok")
