"""Torch restatement of the reference's *simulated* QuantLayer forward (fake-quant in the tensor's own dtype, then
F.linear) — TEST INFRASTRUCTURE ONLY.  It runs on whatever device the tensors live on, so GPU model-level tests can
compare the integer kernels with the reference's simulation on the *same* back end (same attention / LayerNorm /
cuBLAS), isolating the kernels' numerics from cross-back-end fp16 noise.  Pinned bit-exact on CPU against the
reference-generated vectors (tests/test_oracle_golden.py::test_torch_fake_quant_*).

Follows dynamic_quantizer.py:16-45, base_quantizer.py:177-228 ('token'), :129-144 (weights), quant_layer.py:140,178,211.
"""
import torch
import torch.nn.functional as F


def act_fake_quant(x, n_bits=8, exact=False):
    """x [B, n, C]: per-token dynamic asymmetric fake-quant with batch-pooled statistics.
    exact=True returns the dequantised tensor in fp32 WITHOUT the reference's rounding of (q - zp) * delta to fp16
    (same integer codes): the quantity the integer GEMM represents exactly."""
    B, n, C = x.shape
    rows = x.permute(1, 0, 2).reshape(n, -1)
    mn = rows.min(dim=-1)[0]
    mn = torch.where(mn > 0, torch.zeros_like(mn), mn)
    mx = rows.max(dim=-1)[0]
    mx = torch.where(mx < 0, torch.zeros_like(mx), mx)
    levels = 2 ** n_bits
    delta = (mx - mn) / (levels - 1)
    if delta.min() < 1e-6:
        delta = torch.full_like(delta, 1e-6)
    zp = torch.round(-mn / delta)
    delta, zp = delta.reshape(1, n, 1), zp.reshape(1, n, 1)
    q = torch.clamp(torch.round(x / delta) + zp, 0, levels - 1)
    if exact:
        return (q.float() - zp.float()) * delta.float()
    return (q - zp) * delta


def weight_fake_quant(w, delta, zp, n_bits=8, exact=False):
    d, z = delta.reshape(-1, 1).to(w.dtype), zp.reshape(-1, 1).to(w.dtype)
    q = torch.clamp(torch.round(w / d) + z, 0, 2 ** n_bits - 1)
    if exact:
        return (q.float() - z.float()) * d.float()
    return (q - z) * d


def quant_linear_fake(x, w, b, wdelta, wzp, w_bits=8, a_bits=8, smooth=None, exact=False):
    """x [B, n, C] (already in the layer's statistics view) -> [B, n, N].
    exact=False: the reference's simulation op for op (operands rounded to the tensor dtype, F.linear in that dtype).
    exact=True : same integer codes, un-rounded dequantised operands, fp32 matmul, one rounding of the result — the
                 simulation without its own fp16 operand noise (what a real integer kernel computes)."""
    if smooth is not None:
        x = x / smooth
        w = w * smooth
    xh, wh = act_fake_quant(x, a_bits, exact), weight_fake_quant(w, wdelta, wzp, w_bits, exact)
    if exact:
        return F.linear(xh, wh, None if b is None else b.float()).to(x.dtype)
    return F.linear(xh, wh, b)
