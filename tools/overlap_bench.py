"""Overlapped schedule (policy mode 2: quantiser warpgroups inside the persistent GEMM) against quantise pass + GEMM, per shape,
in a replayed CUDA graph of 20 calls with a 256 MB L2 flush in front of every call pair."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from viditq_b200 import ops  # noqa: E402

dev, K, REP = "cuda", 1152, 20
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def weight(N):
    w = (torch.randn(N, K, device=dev) * 0.03).half()
    mn, mx = w.float().min(1)[0].clamp(max=0), w.float().max(1)[0].clamp(min=0)
    d = ((mx - mn) / 255).half()
    return ops.prep_weight(w, d, torch.round(-mn / d.float()).half(), bias=torch.zeros(N, device=dev).half())


def graph_us(fn):
    fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP):
            flush.zero_()
            fn()
    gf = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gf):
        for _ in range(REP):
            flush.zero_()
    best = []
    for gr in (g, gf):
        for _ in range(2):
            gr.replay()
        torch.cuda.synchronize()
        b = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            b = min(b, e0.elapsed_time(e1))
        best.append(b)
    return (best[0] - best[1]) * 1e3 / REP


print("| rows | N | epilogue | LN | overlapped us | quant + GEMM us |")
print("|---|---|---|---|---|---|")
for (rows, N, epi, ln) in [(32768, 3456, 0, True), (32768, 4608, 0, True), (32768, 3456, 0, False), (32768, 1152, 2, False),
                           (32768, 1152, 0, False), (16384, 3456, 0, True), (16384, 1152, 2, False), (4096, 3456, 0, True),
                           (2048, 3456, 0, True)]:
    x = torch.randn(1, rows, K, device=dev).half()
    pw = weight(N)
    shift = (torch.randn(1, K, device=dev) * 0.1).half() if ln else None
    scale = (torch.randn(1, K, device=dev) * 0.1).half() if ln else None
    res = torch.randn(rows, N, device=dev).half() if epi == 2 else None
    gate = torch.randn(1, N, device=dev).half() if epi == 2 else None
    out = torch.empty(rows, N, device=dev, dtype=torch.float16)
    a_buf = ops._alloc_act(1, rows, K, x.device)

    def one():
        ops.linear_w8a8(x, pw, ln=(shift, scale) if ln else None, epi=epi, res=res, gate=gate, rows_per_gate=rows if epi == 2 else 0,
                        out=out)

    def two():
        if ln:
            a, _ = ops.ln_modulate_act_quant(x, shift, scale, out=a_buf)
        else:
            a = ops.act_quant(x, out=a_buf)
        ops.gemm_w8a8(a, pw, epi=epi, res=res, gate=gate, rows_per_gate=rows if epi == 2 else 0, out=out)
    ops.set_linear_fused_policy(2)
    t1 = graph_us(one)
    ops.set_linear_fused_policy(0)
    t2 = graph_us(two)
    print(f"| {rows} | {N} | {['bias', 'gelu', 'gate+res'][epi]} | {'y' if ln else 'n'} | {t1:.1f} | {t2:.1f} |", flush=True)
