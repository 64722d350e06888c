import os, sys, torch
sys.path.insert(0, "/root/repo")
from viditq_b200 import ops
M, C = 32768, 1152
torch.manual_seed(0)
x = (torch.randn(1, M, C, device="cuda") * 1.5).half()
shift = (torch.randn(2, C, device="cuda") * 0.1).half(); scale = (torch.randn(2, C, device="cuda") * 0.1).half()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, n=20):
    for _ in range(3): fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for s, e in ev:
        flush.zero_(); s.record(); fn(); e.record()
    torch.cuda.synchronize()
    return sorted(s.elapsed_time(e) for s, e in ev)[n // 2] * 1e3
a = ops.act_quant(x)
print(os.environ.get("VQ_AQ_NOPF", "0"), "plain %.1f us" % t(lambda: ops.act_quant(x, out=a)), "LN %.1f us" % t(lambda: ops.ln_modulate_act_quant(x, shift, scale, rows_per_mod=M // 2, out=a)))
