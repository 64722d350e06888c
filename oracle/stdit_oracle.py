"""CPU oracle of one W8A8 STDiT block (reference t2v/opensora/models/stdit/stdit.py:96-133 with every nn.Linear
replaced by the fake-quant QuantLayer family, qdiff/models/stdit_quant_layer.py).  TEST INFRASTRUCTURE ONLY — used by
bench.py's cpu_baseline / `--impl reference` legs (the Python reference cannot travel to the GPU box) and by tests.

numpy restatement; the quantised linears go through oracle.qdiff_oracle (pinned bit-exact against the reference).
Attention is plain softmax attention in fp32 (blocks.py:179-187 eager branch; 72^-0.5 scale).
"""
import numpy as np

from . import qdiff_oracle as O

F16, F32 = np.float16, np.float32


def make_block_params(seed, C=1152, mlp_ratio=4):
    """Seeded synthetic weights + min-max weight quantiser parameters for the 13 quantised linears of a block."""
    rng = np.random.default_rng(seed)
    P = {}

    def lin(name, n_out, n_in):
        bound = np.sqrt(6.0 / (n_in + n_out))
        w = rng.uniform(-bound, bound, size=(n_out, n_in)).astype(F32)
        b = (rng.standard_normal(n_out) * 0.02).astype(F16)
        d, z = O.weight_init_params(w, 8)
        P[name] = dict(w=w.astype(F16), b=b, d=d.astype(F16), z=z.astype(F16))

    for a in ("attn", "attn_temp"):
        for l in ("q", "k", "v", "proj"):
            lin(f"{a}.{l}", C, C)
    lin("cross_attn.q_linear", C, C)
    lin("cross_attn.kv_linear", 2 * C, C)
    lin("cross_attn.proj", C, C)
    lin("mlp.fc1", mlp_ratio * C, C)
    lin("mlp.fc2", C, mlp_ratio * C)
    P["scale_shift_table"] = (rng.standard_normal((6, C)) / np.sqrt(C)).astype(F16)
    return P


def qlinear(x16, p):
    """QuantLayer forward: dynamic per-token act fake-quant -> static per-channel weight fake-quant -> fp16 linear."""
    return O.quant_linear_fake(x16, p["w"], p["b"], p["d"], p["z"], 8, 8)


def _softmax_attention(q, k, v, scale):
    """q [n, Lq, H, D], k/v [n, Lk, H, D] -> [n, Lq, H, D], fp32 math, fp16 result."""
    qf = q.astype(F32).transpose(0, 2, 1, 3)
    kf = k.astype(F32).transpose(0, 2, 3, 1)
    vf = v.astype(F32).transpose(0, 2, 1, 3)
    s = (qf @ kf) * F32(scale)
    s = s - s.max(axis=-1, keepdims=True)
    p = np.exp(s)
    p /= p.sum(axis=-1, keepdims=True)
    return (p @ vf).transpose(0, 2, 1, 3).astype(F16)


def _h16(x):
    return np.asarray(x, dtype=F32).astype(F16)


def stdit_block(x16, y16, t0_16, P, T, S, y_lens, H=16):
    """x16 [B, T*S, C]; y16 [1, sum(y_lens), C]; t0_16 [B, 6*C].  Returns the block output, fp16."""
    B, N, C = x16.shape
    D = C // H
    mod = _h16(P["scale_shift_table"].astype(F32)[None] + t0_16.astype(F32).reshape(B, 6, C))
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = (mod[:, i] for i in range(6))

    def gated(x, g, y):
        return _h16(x.astype(F32) + _h16(g.astype(F32)[:, None, :] * y.astype(F32)).astype(F32))

    # spatial
    xm = O.ln_modulate(x16, shift_msa, scale_msa)
    xs = xm.reshape(B * T, S, C)
    pv = xs.reshape(B, T * S, C)            # stdit_quant_layer.py:70 statistics view
    q, k, v = (qlinear(pv, P[f"attn.{n}"]).reshape(B * T, S, H, D) for n in ("q", "k", "v"))
    o = _softmax_attention(q, k, v, D ** -0.5).reshape(B, N, C)
    x16 = gated(x16, gate_msa, qlinear(o, P["attn.proj"]))
    # temporal
    xt = x16.reshape(B, T, S, C).transpose(0, 2, 1, 3).reshape(B, S * T, C)
    q, k, v = (qlinear(xt, P[f"attn_temp.{n}"]).reshape(B * S, T, H, D) for n in ("q", "k", "v"))
    o = _softmax_attention(q, k, v, D ** -0.5).reshape(B, S * T, C)
    o = qlinear(o, P["attn_temp.proj"]).reshape(B, S, T, C).transpose(0, 2, 1, 3).reshape(B, N, C)
    x16 = gated(x16, gate_msa, o)
    # cross
    q = qlinear(x16, P["cross_attn.q_linear"]).reshape(B, N, H, D)
    kv = qlinear(y16, P["cross_attn.kv_linear"]).reshape(-1, 2, H, D)
    outs, off = [], 0
    for b in range(B):
        L = y_lens[b]
        outs.append(_softmax_attention(q[b:b + 1], kv[None, off:off + L, 0], kv[None, off:off + L, 1], D ** -0.5))
        off += L
    o = np.concatenate(outs, 0).reshape(B, N, C)
    x16 = _h16(x16.astype(F32) + qlinear(o, P["cross_attn.proj"]).astype(F32))
    # mlp
    xm = O.ln_modulate(x16, shift_mlp, scale_mlp)
    h = _h16(O.gelu_tanh(qlinear(xm, P["mlp.fc1"]).astype(F32)))
    x16 = gated(x16, gate_mlp, qlinear(h, P["mlp.fc2"]))
    return x16
