"""CPU oracle for the ViDiT-Q quantised-linear hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the
product path (vidit-q_b200/) never does and fails loudly when its CUDA library is missing.

This is a numpy restatement of the reference's arithmetic (thu-nics/ViDiT-Q @ 44126bd), each function citing the
file:line it follows.  The reference runs the model and the quantiser buffers in fp16 (t2v/scripts/quant_txt2video.py:207
`qnn.to(dtype)`), so every torch op on a half tensor is restated as "fp32 compute, one round-to-nearest-even to fp16"
(`_h`).  Parity pinning: tests/golden/*.npz hold inputs/outputs produced by the *unmodified reference classes* imported
from /root/reference (generator: tests/golden/make_golden.py); tests/test_oracle_golden.py checks this file against
them bit-for-bit (codes, delta, zero-point, fake-quant tensors) — the reference itself ships no tests or golden vectors
(SURVEY.md §4), so that is the only pin that exists.
"""
import numpy as np

F16 = np.float16
F32 = np.float32


def _h(x):
    """Round an fp32 array to fp16 and hold it in fp32 (the result of one torch op on a half tensor)."""
    with np.errstate(over="ignore", invalid="ignore"):
        return np.asarray(x, dtype=F32).astype(F16).astype(F32)


def as_f32(x16):
    return np.asarray(x16).astype(F32)


# ----------------------------------------------------------------------------------------------------------------------
# a1: DynamicActQuantizer (per-token, batch-pooled statistics)
# ----------------------------------------------------------------------------------------------------------------------
def token_quant_params(x16, n_bits=8):
    """base_quantizer.py:177-228 ('token' branch, scale_method 'min_max', sym False).

    x16: fp16 [B, n_token, C].  Rows are tokens, statistics pooled over batch x channel (quirk Q1, :184-185), the range
    always contains zero (:191-194), delta = (max-min)/(2^b-1) (:219), the eps quirk fills *all* rows (:220-223),
    zero_point = round(-min/delta) (:228).  Returns (delta [n], zp [n], degenerate flag), fp16 values in fp32.
    """
    x = as_f32(x16)
    B, n, C = x.shape
    rows = x.transpose(1, 0, 2).reshape(n, B * C)
    mn = np.minimum(rows.min(axis=-1), F32(0))
    mx = np.maximum(rows.max(axis=-1), F32(0))
    qmax = F32(2 ** n_bits - 1)
    rng = _h(mx - mn)
    delta = _h(rng / qmax)
    degenerate = bool(delta.min() < F32(1e-6))
    if degenerate:
        delta = np.full_like(delta, _h(F32(1e-6)))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        zp = np.rint(_h(-mn / delta))
    return delta, zp, degenerate


def quant_codes(x16, delta, zp, n_bits=8):
    """dynamic_quantizer.py:36-37 / base_quantizer.py:134-140: clamp(round(x/delta) + zp, 0, 2^b - 1).

    delta / zp must broadcast against x16.  Returns integer codes as fp32 (NaN possible only in the degenerate case).
    """
    qmax = F32(2 ** n_bits - 1)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        y = _h(as_f32(x16) / delta)
        q = np.rint(y) + zp
    return np.clip(q, F32(0), qmax)


def dequant(q, delta, zp):
    """dynamic_quantizer.py:44 / base_quantizer.py:143: (x_quant - zero_point) * delta, fp16 result."""
    return _h((q - zp) * delta).astype(F16)


def dynamic_act_quant(x16, n_bits=8, smooth16=None):
    """Full a1: returns dict(codes u8 [B,n,C], delta [n], zp [n], rowsum i32 [B,n], xhat fp16 [B,n,C], degenerate).

    smooth16 (fp16 [C]) applies `input = input / channel_wise_scale` first (quant_layer.py:140).
    """
    x = as_f32(x16)
    if smooth16 is not None:
        x = _h(x / as_f32(smooth16))
    delta, zp, deg = token_quant_params(x.astype(F16), n_bits)
    d3, z3 = delta[None, :, None], zp[None, :, None]
    q = quant_codes(x.astype(F16), d3, z3, n_bits)
    out = dict(delta=delta, zp=zp, degenerate=deg, xhat=dequant(q, d3, z3))
    if not deg:
        out["codes"] = q.astype(np.uint8)
        out["rowsum"] = q.astype(np.int64).sum(axis=-1).astype(np.int32)
    return out


def static_act_quant(x16, delta16, zp16, n_bits=8, smooth16=None):
    """BaseQuantizer.forward with init_done (base_quantizer.py:129-143) for a calibrated ActQuantizer: delta16 / zp16 are
    the checkpoint's fp16 buffers, 1 element (per_group False: w8a8_naive.yaml) or n_token elements (static per-token,
    broadcast as [1, n_token, 1]).  Returns dict(codes u8, rowsum i32, xhat fp16) for x16 fp16 [B, n, C]."""
    x = as_f32(x16)
    if smooth16 is not None:
        x = _h(x / as_f32(smooth16))
    d, z = as_f32(delta16).reshape(-1), as_f32(zp16).reshape(-1)
    if d.size > 1:
        d, z = d[None, :, None], z[None, :, None]
    q = quant_codes(x.astype(F16), d, z, n_bits)
    return dict(codes=q.astype(np.uint8), rowsum=q.astype(np.int64).sum(axis=-1).astype(np.int32),
                xhat=dequant(q, d, z))


# ----------------------------------------------------------------------------------------------------------------------
# a2: WeightQuantizer with static per-output-channel parameters
# ----------------------------------------------------------------------------------------------------------------------
def weight_init_params(w16, n_bits=8):
    """base_quantizer.py:166-228 ('channel' branch, channel_dim 0): what ptq.py writes into ckpt.pth (delta, zero_point
    of shape [N,1]).  Computed in the dtype of w (PTQ runs fp32 weights -> pass float32 to mirror it)."""
    w = np.asarray(w16)
    f = (lambda v: v.astype(F32)) if w.dtype == np.float32 else _h
    wf = w.astype(F32)
    mn = np.minimum(wf.min(axis=-1), F32(0))
    mx = np.maximum(wf.max(axis=-1), F32(0))
    qmax = F32(2 ** n_bits - 1)
    delta = f(f(mx - mn) / qmax)
    if delta.min() < F32(1e-6):
        delta = np.full_like(delta, f(np.asarray(F32(1e-6))))
    zp = np.rint(f(-mn / delta))
    return delta, zp


def weight_quant(w16, delta, zp, n_bits=8, smooth16=None):
    """base_quantizer.py:129-144 on an fp16 weight [N,K] with fp16 delta/zp [N]; optional `weight * channel_wise_scale`
    (quant_layer.py:178).  Returns dict(codes u8, colsum i32 [N], what fp16 [N,K])."""
    w = as_f32(w16)
    if smooth16 is not None:
        w = _h(w * as_f32(smooth16)[None, :])
    d2, z2 = as_f32(delta).reshape(-1, 1), as_f32(zp).reshape(-1, 1)
    q = quant_codes(w.astype(F16), d2, z2, n_bits)
    return dict(codes=q.astype(np.uint8), colsum=q.astype(np.int64).sum(axis=-1).astype(np.int32),
                what=dequant(q, d2, z2))


def smooth_channel_scale(act_scale16, w16, alpha):
    """quant_layer.py:137 (momentum type): act_scale.pow(alpha) / weight.abs().max(dim=0)[0].pow(1 - alpha), fp16 ops.
    torch casts a Python-scalar exponent to the tensor dtype (fp16) before powf, on CPU and CUDA alike."""
    a = _h(np.power(as_f32(act_scale16), F32(F16(alpha))))
    wmax = np.abs(as_f32(w16)).max(axis=0)
    b = _h(np.power(wmax, F32(F16(1.0 - alpha))))
    return _h(a / b).astype(F16)


# ----------------------------------------------------------------------------------------------------------------------
# a3-a7: QuantLayer-family forward
# ----------------------------------------------------------------------------------------------------------------------
def linear_f16(xhat16, what16, bias16=None):
    """F.linear on fp16 operands (quant_layer.py:211): fp32 accumulate, one rounding of the result to fp16."""
    acc = as_f32(xhat16) @ as_f32(what16).T
    if bias16 is not None:
        acc = acc + as_f32(bias16)
    return _h(acc).astype(F16)


def quant_linear_fake(x16, w16, bias16, wdelta, wzp, w_bits=8, a_bits=8, smooth16=None):
    """The reference's simulated path: act fake-quant -> weight fake-quant -> F.linear, x16 [B,n,C] -> [B,n,N]."""
    a = dynamic_act_quant(x16, a_bits, smooth16)
    wq = weight_quant(w16, wdelta, wzp, w_bits, smooth16)
    B, n, C = x16.shape
    return linear_f16(a["xhat"].reshape(B * n, C), wq["what"], bias16).reshape(B, n, -1)


def gelu_tanh(x):
    x = np.asarray(x, dtype=np.float64)
    return 0.5 * x * (1.0 + np.tanh(0.7978845608028654 * (x + 0.044715 * x ** 3)))


def quant_linear_int(codes, delta, zp, rowsum, wcodes, wdelta, wzp, bias, epi="bias", res16=None, gate16=None):
    """The integer decomposition the CUDA GEMM evaluates (same operation order as vq_gemm_w8a8's epilogue):
        t = sum_k xq wq - zx * c1 - rowsum * zw ;  y = h(fma(float(t), dx * dw, bias))
    codes [B,n,K] u8, delta/zp [n], rowsum [B,n], wcodes [N,K].  Returns fp16 [B,n,N]."""
    B, n, K = codes.shape
    acc = codes.reshape(B * n, K).astype(np.int64) @ wcodes.astype(np.int64).T
    zw = np.rint(as_f32(wzp)).astype(np.int64).reshape(-1)
    c1 = wcodes.astype(np.int64).sum(axis=-1) - K * zw
    zx = np.tile(np.rint(zp).astype(np.int64), B)
    t = acc - zx[:, None] * c1[None, :] - rowsum.reshape(-1).astype(np.int64)[:, None] * zw[None, :]
    assert np.abs(t).max() < 2 ** 31 and np.abs(acc).max() < 2 ** 31
    tf = t.astype(F32).astype(np.float64)
    s = (np.tile(delta.astype(F32), B)[:, None] * as_f32(wdelta).reshape(1, -1)).astype(np.float64)  # exact in fp32
    b = np.zeros(wcodes.shape[0], F32) if bias is None else as_f32(bias)
    y = _h((tf * s + b.astype(np.float64)).astype(F32))
    if epi == "gelu_tanh":
        y = _h(gelu_tanh(y).astype(F32))
    elif epi == "gate_residual":
        g = as_f32(gate16).reshape(B, 1, -1)
        y = _h(as_f32(res16).reshape(B, n, -1) + _h(g * y.reshape(B, n, -1))).reshape(B * n, -1)
    return y.astype(F16).reshape(B, n, -1)


# ----------------------------------------------------------------------------------------------------------------------
# a9 glue: LayerNorm + t2i_modulate (stdit.py:104,125; blocks.py:51)
# ----------------------------------------------------------------------------------------------------------------------
def ln_modulate(x16, shift16, scale16, eps=1e-6):
    """nn.LayerNorm(C, eps=1e-6, elementwise_affine=False) on a half tensor, then x * (1 + scale) + shift in fp16 ops.
    x16 [B,n,C]; shift16/scale16 [B,C]."""
    x = as_f32(x16).astype(np.float64)
    mean = x.mean(axis=-1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
    ln = _h(((x - mean) / np.sqrt(var + eps)).astype(F32))
    one_plus = _h(F32(1) + as_f32(scale16))[:, None, :]
    return _h(_h(ln * one_plus) + as_f32(shift16)[:, None, :]).astype(F16)
