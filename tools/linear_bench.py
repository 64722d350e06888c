"""Fused one-launch QuantLinear (vq_linear_fused_kernel) against quantise pass + GEMM, per shape, inside a replayed CUDA graph
of 40 back-to-back calls (the regime of the fused schedules) and as eager launches (the regime of the hook API).

    VQ_LINEAR_FUSED=1 python tools/linear_bench.py        (the env forces the fused kernel at every supported shape)
"""
import os
import sys
import time

os.environ.setdefault("VQ_LINEAR_FUSED", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from viditq_b200 import ops  # noqa: E402

dev = "cuda"
K = 1152
REP = 40


def weight(N, n_bits=8):
    w = (torch.randn(N, K, device=dev) * 0.03).half()
    mn, mx = w.float().min(1)[0].clamp(max=0), w.float().max(1)[0].clamp(min=0)
    d = ((mx - mn) / (2 ** n_bits - 1)).half()
    return ops.prep_weight(w, d, torch.round(-mn / d.float()).half(), n_bits=n_bits, bias=torch.zeros(N, device=dev).half())


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
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3 / REP


def eager_us(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e6 / 200


print("| G x rows | N | epilogue | LN | W | graph: fused us | graph: quant+GEMM us | eager: fused us | eager: quant+GEMM us |")
print("|---|---|---|---|---|---|---|---|---|")
for (G, rows, N, epi, ln, wb) in [(1, 109, 2304, 0, False, 8), (1, 218, 2304, 0, False, 8), (1, 2048, 1152, 2, False, 8),
                                  (1, 2048, 3456, 0, True, 8), (1, 2048, 4608, 0, True, 8), (2, 1024, 1152, 2, False, 8),
                                  (2, 1024, 3456, 0, True, 8), (2, 1024, 4608, 0, True, 8), (2, 1024, 3456, 0, True, 4),
                                  (1, 4096, 1152, 2, False, 8), (1, 4096, 3456, 0, True, 8), (1, 8192, 1152, 2, False, 8),
                                  (1, 8192, 3456, 0, True, 8), (1, 16384, 1152, 2, False, 8), (1, 16384, 3456, 0, True, 8),
                                  (1, 16384, 4608, 0, True, 8), (1, 32768, 1152, 2, False, 8), (1, 32768, 3456, 0, True, 8)]:
    if ops.linear_launch_count(G, rows, K) != 1:
        print(f"| {G} x {rows} | {N} | - | - | - | unsupported | | | |")
        continue
    M = G * rows
    x = torch.randn(G, rows, K, device=dev).half()
    pw = weight(N, wb)
    if wb == 4:
        ops.pack_u4(pw)
    shift = (torch.randn(G, K, device=dev) * 0.1).half() if ln else None
    scale = (torch.randn(G, K, device=dev) * 0.1).half() if ln else None
    res = torch.randn(M, N, device=dev).half() if epi == 2 else None
    gate = torch.randn(1, N, device=dev).half() if epi == 2 else None
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    a_buf = ops._alloc_act(G, rows, K, x.device)

    def fused():
        ops.linear_w8a8(x, pw, ln=(shift, scale) if ln else None, epi=epi, res=res, gate=gate, rows_per_gate=M if epi == 2 else 0,
                        out=out)

    def two():
        if ln:
            a, _ = ops.ln_modulate_act_quant(x, shift, scale, out=a_buf)
        else:
            a = ops.act_quant(x, out=a_buf)
        ops.gemm_w8a8(a, pw, epi=epi, res=res, gate=gate, rows_per_gate=M if epi == 2 else 0, out=out)
    print(f"| {G} x {rows} | {N} | {['bias', 'gelu', 'gate+res'][epi]} | {'y' if ln else 'n'} | {wb} | {graph_us(fused):.1f} | "
          f"{graph_us(two):.1f} | {eager_us(fused):.1f} | {eager_us(two):.1f} |", flush=True)
