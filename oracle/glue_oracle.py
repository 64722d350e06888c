"""CPU oracle of the fp16 glue kernels around the quantised linears: softmax attention and the patch embedding.
TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline): nothing under vidit-q_b200/ imports it.

Follows the reference's in-tree arithmetic:
  * attention        t2v/opensora/models/layers/blocks.py:161-187 (eager branch of Attention.forward: heads split as
                     "B N (3 H D)", q * head_dim^-0.5, scores @ softmax in fp32, probabilities @ v) and the block-diagonal
                     cross attention of blocks.py:292-310 (per sample: image tokens attend to that sample's prompt rows);
  * patch embedding  blocks.py:91-110 (PatchEmbed3D: Conv3d with kernel = stride = patch, "B C T H W -> B (T H W) C") and the
                     position-embedding add of stdit.py:255-258.
Pinned against outputs of the unmodified reference classes by tests/golden/glue_golden.npz (tests/golden/make_golden_glue.py).
"""
import numpy as np

F16, F32 = np.float16, np.float32


def attention(q, k, v, num_heads, scale=None):
    """q: [B, Nq, C], k / v: [B, Nk, C] (C = H * D) -> [B, Nq, C]; fp32 softmax attention per head."""
    B, Nq, C = q.shape
    D = C // num_heads
    scale = D ** -0.5 if scale is None else scale
    qh = q.astype(F32).reshape(B, Nq, num_heads, D).transpose(0, 2, 1, 3) * F32(scale)
    kh = k.astype(F32).reshape(B, -1, num_heads, D).transpose(0, 2, 1, 3)
    vh = v.astype(F32).reshape(B, -1, num_heads, D).transpose(0, 2, 1, 3)
    s = qh @ kh.transpose(0, 1, 3, 2)
    p = np.exp(s - s.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    return (p @ vh).transpose(0, 2, 1, 3).reshape(B, Nq, C)


def cross_attention(q, kv, lens, num_heads):
    """q: [B, N, C]; kv: [sum(lens), 2C] (k | v rows of every sample's prompt, packed) -> [B, N, C]."""
    B, N, C = q.shape
    out, off = [], 0
    for b in range(B):
        L = int(lens[b])
        out.append(attention(q[b:b + 1], kv[None, off:off + L, :C], kv[None, off:off + L, C:], num_heads))
        off += L
    return np.concatenate(out, 0)


def patch_embed(latent, weight, bias, pos, patch_hw, fp16_graph=True):
    """latent [B, Cin, T, H, W]; weight [C, Cin, 1, ph, pw]; bias [C]; pos [S, C] or None -> [B, T*S, C].
    fp16_graph=True reproduces the rounding points of the reference's half-precision graph: latent and weights are fp16
    values, the convolution accumulates in fp32 and rounds (with the bias) to fp16, the position embedding is a
    separate fp16 add."""
    B, Cin, T, H, W = latent.shape
    ph, pw = patch_hw
    C = weight.shape[0]
    gh, gw = H // ph, W // pw
    x = latent.astype(F16).astype(F32) if fp16_graph else latent.astype(F32)
    w = weight.reshape(C, Cin * ph * pw).astype(F32)
    # patches [B, T, gh, gw, Cin, ph, pw] -> [B, T*S, Cin*ph*pw]
    p = x.reshape(B, Cin, T, gh, ph, gw, pw).transpose(0, 2, 3, 5, 1, 4, 6).reshape(B, T * gh * gw, Cin * ph * pw)
    y = p @ w.T
    if bias is not None:
        y = y + bias.astype(F32)
    if fp16_graph:
        y = y.astype(F16)
        if pos is not None:
            y = (y.reshape(B, T, gh * gw, C) + pos.astype(F16)[None, None]).astype(F16).reshape(B, T * gh * gw, C)
        return y
    if pos is not None:
        y = (y.reshape(B, T, gh * gw, C) + pos.astype(F32)[None, None]).reshape(B, T * gh * gw, C)
    return y
