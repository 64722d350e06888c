"""TEST INFRASTRUCTURE: CPU stand-ins for the viditq_b200.ops kernel wrappers, built on the oracle's integer form.

`patch_ops(monkeypatch)` swaps the functions of `viditq_b200.ops` that the layer-by-layer (hook-API) path calls for numpy /
torch-CPU restatements, so the HOST logic of viditq_b200.qdiff — QuantLayer state handling, `accelerate()` on an unmodified
reference QuantModel, smooth-quant modes, the checkpoint plumbing — can be exercised in this GPU-less container against
the reference executed on the same CPU.  Never imported by the product path (which has no CPU fallback and raises on CPU
tensors); the CUDA kernels themselves are pinned against the same oracle by the `-m gpu` tests.
"""
import numpy as np
import torch

from oracle import qdiff_oracle as O
from viditq_b200 import ops

F16, F32 = np.float16, np.float32


def _np16(t):
    assert t.dtype == torch.float16 and not t.is_cuda, (t.dtype, t.device)
    return t.detach().contiguous().numpy()


def prep_weight(w, delta, zp, n_bits=8, smooth=None, bias=None):
    r = O.weight_quant(_np16(w), _np16(delta.reshape(-1).to(torch.float16)), _np16(zp.reshape(-1).to(torch.float16)), n_bits,
                       None if smooth is None else _np16(smooth.reshape(-1).to(torch.float16)))
    codes = r["codes"]
    N, K = codes.shape
    zw = np.rint(zp.reshape(-1).float().numpy()).astype(np.int64)
    c1 = codes.astype(np.int64).sum(axis=-1) - K * zw
    dw = delta.reshape(-1).to(torch.float16).float().numpy()
    b = np.zeros(N, F32) if bias is None else bias.reshape(-1).to(torch.float16).float().numpy()
    pw = ops.PreparedWeight(torch.from_numpy(codes), (c1, zw, dw, b), N, K, n_bits)
    return pw


def _gelu16(xn):
    return O._h(O.gelu_tanh(xn.astype(F32)).astype(F32)).astype(F16)


def act_quant(x, n_bits=8, smooth=None, out=None, gelu=False):
    xn = _np16(x)
    G, rows, K = xn.shape
    if gelu:
        xn = _gelu16(xn)
    r = O.dynamic_act_quant(xn, n_bits, None if smooth is None else _np16(smooth))
    if r["degenerate"]:
        raise RuntimeError("degenerate eps row (quirk Q4)")
    return ops.ActCodes(torch.from_numpy(r["codes"].reshape(G * rows, K)), torch.from_numpy(r["delta"].astype(F16)),
                        torch.from_numpy(r["zp"].astype(F16)), torch.from_numpy(r["rowsum"].reshape(-1)), G, rows, K)


def act_quant_static(x, delta, zp, n_bits=8, smooth=None, gelu=False, ln=None, rows_per_mod=None):
    assert ln is None, "the LayerNorm-fused static form belongs to the fused schedule (GPU only)"
    xn = _np16(x)
    if gelu:
        xn = _gelu16(xn)
    K = xn.shape[-1]
    period = delta.numel()
    x3 = xn.reshape(-1, period, K) if period > 1 else xn.reshape(1, -1, K)
    r = O.static_act_quant(x3, _np16(delta), _np16(zp), n_bits, None if smooth is None else _np16(smooth))
    M = x3.shape[0] * x3.shape[1]
    return ops.ActCodes(torch.from_numpy(r["codes"].reshape(M, K)), delta.reshape(-1), zp.reshape(-1),
                        torch.from_numpy(r["rowsum"].reshape(-1)), M // period, period, K)


def col_absmax(x, gelu=False):
    if x.dtype == torch.float32 and not gelu:      # PTQ statistics of an fp32 model (tests/test_ptq_cpu.py)
        return x.abs().max(dim=-2)[0]
    xn = _np16(x)
    if gelu:
        xn = _gelu16(xn)
    return torch.from_numpy(np.abs(xn).max(axis=-2))


def gemm_w8a8(a, w, epi=ops.VQ_EPI_BIAS, res=None, gate=None, rows_per_gate=0, out=None, ldo=None):
    """The epilogue order of vq_gemm_w8a8: t = acc - zx*c1 - rowsum*zw (int32), y = h(fma(float(t), dx*dw, bias)), then the
    GELU / gated-residual variants in fp16 steps.  The accumulator is evaluated in float64 BLAS (exact: < 2^53)."""
    c1, zw, dw, b = w.col
    M = a.G * a.rows
    acc = a.codes.numpy().astype(np.float64) @ w.codes.numpy().astype(np.float64).T
    zx = np.tile(np.rint(a.zp.float().numpy()).astype(np.int64), M // a.rows)
    dx = np.tile(a.delta.float().numpy(), M // a.rows)
    t = acc.astype(np.int64) - zx[:, None] * c1[None, :] - a.rowsum.numpy().astype(np.int64)[:, None] * zw[None, :]
    assert np.abs(t).max() < 2 ** 31
    s = (dx[:, None].astype(F32) * dw[None, :].astype(F32)).astype(np.float64)
    y = O._h((t.astype(F32).astype(np.float64) * s + b[None, :].astype(np.float64)).astype(F32))
    if epi == ops.VQ_EPI_GELU_TANH:
        y = O._h(O.gelu_tanh(y).astype(F32))
    elif epi == ops.VQ_EPI_GATE_RESIDUAL:
        g = np.repeat(gate.float().numpy().reshape(-1, w.N), rows_per_gate, axis=0)[:M]
        y = O._h(res.float().numpy().reshape(M, w.N) + O._h(g * y))
    y = torch.from_numpy(y.astype(F16))
    if out is not None:
        out.copy_(y)
        return out
    return y


def linear_w8a8(x, w, n_bits=8, smooth=None, ln=None, rows_per_mod=None, epi=ops.VQ_EPI_BIAS, res=None, gate=None,
                rows_per_gate=0, out=None, ldo=None):
    assert ln is None, "LN-fused linear is a fused-schedule feature (GPU tests)"
    return gemm_w8a8(act_quant(x, n_bits=n_bits, smooth=smooth), w, epi=epi, res=res, gate=gate, rows_per_gate=rows_per_gate,
                     out=out, ldo=ldo)


def _row_positions(rows, dims, strides):
    idx = np.arange(rows)
    i3 = idx % dims[3]
    q = idx // dims[3]
    i2 = q % dims[2]
    q = q // dims[2]
    i1 = q % dims[1]
    i0 = q // dims[1]
    return i0 * strides[0] + i1 * strides[1] + i2 * strides[2] + i3 * strides[3]


def pack_rows(a, dims, strides):
    """numpy restatement of vq_row_pack (pack): codes + {delta, zp, rowsum} tail, rows scattered to the permuted position."""
    rows, K = a.codes.shape[0], a.K
    out = np.zeros((rows, K + 16), np.uint8)
    pos = _row_positions(rows, dims, strides)
    out[pos, :K] = a.codes.numpy()
    tail = np.zeros((rows, 4), np.int32)
    tail[:, 0] = (a.delta.numpy().view(np.uint16).astype(np.uint32) | (a.zp.numpy().view(np.uint16).astype(np.uint32) << 16)).view(np.int32)
    tail[:, 1] = a.rowsum.numpy()
    out[pos, K:] = tail.view(np.uint8).reshape(rows, 16)
    return torch.from_numpy(out)


def unpack_rows(buf, K, dims, strides):
    b = buf.numpy()
    rows = b.shape[0]
    pos = _row_positions(rows, dims, strides)
    codes = np.zeros((rows, K), np.uint8)
    codes[pos] = b[:, :K]
    tail = np.ascontiguousarray(b[:, K:]).view(np.int32).reshape(rows, 4)
    dz = tail[:, 0].view(np.uint32)
    delta = np.zeros(rows, np.float16)
    zp = np.zeros(rows, np.float16)
    rowsum = np.zeros(rows, np.int32)
    delta[pos] = (dz & 0xffff).astype(np.uint16).view(np.float16)
    zp[pos] = (dz >> 16).astype(np.uint16).view(np.float16)
    rowsum[pos] = tail[:, 1]
    return ops.ActCodes(torch.from_numpy(codes), torch.from_numpy(delta), torch.from_numpy(zp), torch.from_numpy(rowsum), 1, rows, K)


def patch_ops(monkeypatch):
    for name, fn in (("prep_weight", prep_weight), ("act_quant", act_quant), ("act_quant_static", act_quant_static),
                     ("col_absmax", col_absmax), ("gemm_w8a8", gemm_w8a8), ("linear_w8a8", linear_w8a8),
                     ("pack_u4", lambda w: w)):
        monkeypatch.setattr(ops, name, fn)
