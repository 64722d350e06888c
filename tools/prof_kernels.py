"""Launch each hot kernel a few times at STDiT 16x512x512 shapes (for ncu captures and CUDA-event timing)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viditq_b200 import ops

torch.manual_seed(0)
M, C = 16384, 1152
dev = "cuda"
x = (torch.randn(1, M, C, device=dev) * 2).half()
h = torch.randn(1, M, 4 * C, device=dev).half()
shift = (torch.randn(1, C, device=dev) * 0.1).half()
scale = (torch.randn(1, C, device=dev) * 0.1).half()


def mk_w(N, K):
    w = (torch.randn(N, K, device=dev) * 0.02).half()
    mn, mx = w.float().min(1)[0].clamp(max=0), w.float().max(1)[0].clamp(min=0)
    d = ((mx - mn) / 255)
    return ops.prep_weight(w, d.half(), torch.round(-mn / d).half(), bias=torch.zeros(N, device=dev).half())


w_sq, w_qkv, w_fc1, w_fc2 = mk_w(C, C), mk_w(3 * C, C), mk_w(4 * C, C), mk_w(C, 4 * C)
gate = torch.randn(1, C, device=dev).half()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(name, fn, algo_bytes=None, ops_=None, iters=20):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for s, e in ev:
        flush.zero_()          # L2 flush between timed iterations
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    ms = sorted(s.elapsed_time(e) for s, e in ev)[iters // 2]
    msg = f"{name:34s} {ms * 1e3:8.1f} us"
    if algo_bytes:
        msg += f"  {algo_bytes / ms / 1e6:8.0f} GB/s"
    if ops_:
        msg += f"  {ops_ / ms / 1e9:8.0f} TOP/s"
    print(msg, flush=True)


a = ops.act_quant(x)
a4 = ops.act_quant(h)
qkv = ops.gemm_w8a8(a, w_qkv)
xr = x.view(M, C)
timeit("act_quant K=1152", lambda: ops.act_quant(x, out=a), algo_bytes=M * C * 3 + 8 * M)
timeit("ln_modulate_act_quant K=1152", lambda: ops.ln_modulate_act_quant(x, shift, scale, out=a), algo_bytes=M * C * 3)
timeit("act_quant K=4608", lambda: ops.act_quant(h, out=a4), algo_bytes=M * 4 * C * 3)
timeit("gelu + act_quant K=4608", lambda: ops.act_quant(h, out=a4, gelu=True), algo_bytes=M * 4 * C * 3)
o1 = torch.empty(M, C, device=dev, dtype=torch.float16)
o3 = torch.empty(M, 3 * C, device=dev, dtype=torch.float16)
o4 = torch.empty(M, 4 * C, device=dev, dtype=torch.float16)
timeit("gemm N=1152 K=1152 bias", lambda: ops.gemm_w8a8(a, w_sq, out=o1), ops_=2.0 * M * C * C)
timeit("gemm N=1152 K=1152 gate+res", lambda: ops.gemm_w8a8(a, w_sq, epi=ops.VQ_EPI_GATE_RESIDUAL, res=xr, gate=gate,
                                                            rows_per_gate=M, out=o1), ops_=2.0 * M * C * C)
timeit("gemm N=3456 K=1152 bias (qkv)", lambda: ops.gemm_w8a8(a, w_qkv, out=o3), ops_=2.0 * M * C * 3 * C)
timeit("gemm N=4608 K=1152 gelu (fc1)", lambda: ops.gemm_w8a8(a, w_fc1, epi=ops.VQ_EPI_GELU_TANH, out=o4),
       ops_=2.0 * M * C * 4 * C)
timeit("gemm N=1152 K=4608 gate+res (fc2)", lambda: ops.gemm_w8a8(a4, w_fc2, epi=ops.VQ_EPI_GATE_RESIDUAL, res=xr,
                                                                 gate=gate, rows_per_gate=M, out=o1),
       ops_=2.0 * M * C * 4 * C)
timeit("attn_temporal T=16 S=1024", lambda: ops.attn_temporal(qkv, 1, 16, 1024, 16, 72, 72 ** -0.5, out=o1),
       algo_bytes=M * C * 2 * 4)
timeit("attn_spatial (tcgen05) 16 x 1024", lambda: ops.attn_spatial(qkv, 16, 1024, 16, 72, 72 ** -0.5, out=o1),
       ops_=4.0 * 16 * 16 * 1024 * 1024 * 72)
kv = torch.randn(109, 2 * C, device=dev).half()
st, ln = torch.zeros(1, dtype=torch.int32, device=dev), torch.full((1,), 109, dtype=torch.int32, device=dev)
timeit("attn_cross L=109 (tcgen05)", lambda: ops.attn_cross(xr, kv, st, ln, 1, M, 16, 72, 109, 72 ** -0.5, out=o1),
       algo_bytes=M * C * 2 * 2)
z = torch.randn(1, 4, 16, 64, 64, device=dev)
conv_w = (torch.randn(C, 4, 1, 2, 2, device=dev) * 0.1).half()
pos = torch.randn(1024, C, device=dev).half()
timeit("patch_embed + pos_embed", lambda: ops.patch_embed(z, conv_w, shift.view(-1), pos, (2, 2)), algo_bytes=M * C * 2)
import torch.nn.functional as F
q5 = qkv.view(16, 1024, 3, 16, 72)
for name, be in (("cudnn", torch.nn.attention.SDPBackend.CUDNN_ATTENTION),
                 ("flash", torch.nn.attention.SDPBackend.FLASH_ATTENTION)):
    def sp():
        with torch.nn.attention.sdpa_kernel(be):
            o = F.scaled_dot_product_attention(q5[:, :, 0].transpose(1, 2), q5[:, :, 1].transpose(1, 2),
                                               q5[:, :, 2].transpose(1, 2), scale=72 ** -0.5)
        return o.transpose(1, 2).reshape(1, M, C)
    try:
        timeit(f"spatial SDPA {name} (+layout)", sp, ops_=4.0 * 16 * 16 * 1024 * 1024 * 72)
    except Exception as e:  # noqa: BLE001
        print(f"spatial SDPA {name}: {type(e).__name__}: {e}")
# library yardstick for the GEMM shapes: cuBLASLt fp8 (same bytes, same tensor rate as INT8) with row/column scales, fp16 out
try:
    f8 = torch.float8_e4m3fn
    for (n_, k_) in ((C, C), (3 * C, C), (4 * C, C), (C, 4 * C)):
        A8 = torch.randn(M, k_, device=dev).to(f8)
        B8 = torch.randn(n_, k_, device=dev).to(f8)
        sa = torch.rand(M, 1, device=dev) + 0.5
        sb = torch.rand(1, n_, device=dev) + 0.5
        bias8 = torch.randn(n_, device=dev, dtype=torch.float16)
        o8 = torch.empty(M, n_, device=dev, dtype=torch.float16)
        timeit(f"cublasLt fp8 rowwise N={n_} K={k_}",
               lambda: torch._scaled_mm(A8, B8.t(), scale_a=sa, scale_b=sb, bias=bias8, out_dtype=torch.float16, out=o8),
               ops_=2.0 * M * n_ * k_)
except Exception as e:  # noqa: BLE001
    print(f"cublasLt fp8 yardstick: {type(e).__name__}: {str(e)[:200]}")
try:
    for (n_, k_) in ((C, C), (3 * C, C), (4 * C, C), (C, 4 * C)):
        A8 = torch.randint(-8, 8, (M, k_), device=dev, dtype=torch.int8)
        B8 = torch.randint(-8, 8, (k_, n_), device=dev, dtype=torch.int8)
        timeit(f"cublasLt int8->int32 N={n_} K={k_}", lambda: torch._int_mm(A8, B8), ops_=2.0 * M * n_ * k_)
except Exception as e:  # noqa: BLE001
    print(f"cublasLt int8 yardstick: {type(e).__name__}: {str(e)[:200]}")
def sp_only():
    return F.scaled_dot_product_attention(q5[:, :, 0].transpose(1, 2), q5[:, :, 1].transpose(1, 2),
                                          q5[:, :, 2].transpose(1, 2), scale=72 ** -0.5)
timeit("spatial SDPA default (no out copy)", sp_only, ops_=4.0 * 16 * 16 * 1024 * 1024 * 72)
oh_raw = sp_only()
print("sdpa out contiguous [BT,H,S,D]:", oh_raw.is_contiguous(), tuple(oh_raw.shape), tuple(oh_raw.stride()))
oh = oh_raw.contiguous()
ah = ops.act_quant_heads(oh, 1, M, 1024)
timeit("act_quant_heads (head-major in)", lambda: ops.act_quant_heads(oh, 1, M, 1024), algo_bytes=M * C * 3)
try:
    from flash_attn import flash_attn_func
    timeit("spatial flash_attn_func", lambda: flash_attn_func(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2],
                                                              softmax_scale=72 ** -0.5),
           ops_=4.0 * 16 * 16 * 1024 * 1024 * 72)
except Exception as e:  # noqa: BLE001
    print("flash_attn_func:", type(e).__name__, e)
ops.check_status()
