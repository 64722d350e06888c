"""Quick A/B timing of the quantise passes at the stacked-step shapes (M = 32768), L2 flushed between launches.
   VQ_AQ_NOPF / VQ_SEG_OCC select the kernel variants (see vq_quant.cu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from viditq_b200 import ops  # noqa: E402

M, C = 32768, 1152
torch.manual_seed(0)
x = (torch.randn(1, M, C, device="cuda") * 1.5).half()
h = (torch.randn(1, M, 4 * C, device="cuda") * 1.5).half()
shift = (torch.randn(2, C, device="cuda") * 0.1).half()
scale = (torch.randn(2, C, device="cuda") * 0.1).half()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t(fn, n=20):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for s, e in ev:
        flush.zero_()
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    return sorted(s.elapsed_time(e) for s, e in ev)[n // 2] * 1e3


a = ops.act_quant(x)
a4 = ops.act_quant(h)
print("NOPF", os.environ.get("VQ_AQ_NOPF", "3"), "OCC", os.environ.get("VQ_SEG_OCC", "4"),
      "plain %.1f us" % t(lambda: ops.act_quant(x, out=a)),
      "LN %.1f us" % t(lambda: ops.ln_modulate_act_quant(x, shift, scale, rows_per_mod=M // 2, out=a)),
      "gelu K=4608 %.1f us" % t(lambda: ops.act_quant(h, out=a4, gelu=True)),
      "plain K=4608 %.1f us" % t(lambda: ops.act_quant(h, out=a4)))
