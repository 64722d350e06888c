"""CPU restatement of the opt-in INT8 Q/K/V spatial attention (`vq_attn_spatial_i8`) — TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED BY CONSTRUCTION: the reference never quantises Q / K / V or the probabilities — the hooks are commented
out (qdiff/models/quant_block.py:617-623, :630-632) and STDiT / PixArt attention runs flash-attn on the fp16 linear
outputs (t2v/opensora/models/layers/blocks.py:169-188).  There is therefore no reference output to pin this against;
the scheme below is this repo's own (opt-in, own tolerance), and the tests state two tolerances:
kernel vs this restatement (same integers, fp32 softmax), and this scheme vs the fp16 attention the reference computes.

Scheme (per sequence of S tokens, per head, head_dim 72):
  Q8 = rint(Q * (1 / sq)), sq = max|Q[token, head, :]| / 127                  per token, per head (fp32: one reciprocal
                                                                              per scale, then products — K8, V8 alike)
  K' = K - mean over the sequence's tokens (softmax is invariant to it);
  K8 = rint(K' / sk), sk = max|K'[64-token block, head, :]| / 127             per 64-key block, per head
  V8 = rint(V / sv), sv = max|V[:, head, dim]| / 127 over the sequence        per channel
  S  = (Q8 K8^T) * sq * sk * scale      integer dot products, exact
  P8 = rint(255 * exp(S - m)) as u8, m = the exact row maximum (the kernel takes it in a first pass over the key tiles)
  O  = (sum_k P8 V8) * sv / sum_k (255 * exp(S - m))                          integer accumulation, exact; the normaliser is
                                                                              the un-rounded sum (the rounded bytes' sum would
                                                                              drop the small probabilities from it: 1.6x the error)
"""
import torch

BLOCK_K = 64


def quantise_qkv(qkv, n_seq, S, H, D=72, kmean=None):
    """qkv [n_seq * S, 3 * H * D] fp16/fp32 -> dict of integer codes and scales (fp32 arithmetic, the kernel's formulas).
    kmean [n_seq, H, D]: use this mean of K instead of computing it (a sum over 1024 tokens depends on the summation
    order in its last bits; the code tests feed the kernel's own mean so that the codes compare bit for bit)."""
    x = qkv.float().reshape(n_seq, S, 3, H, D)
    q, k, v = x[:, :, 0], x[:, :, 1], x[:, :, 2]
    kmean = k.mean(dim=1, keepdim=True) if kmean is None else kmean.float().reshape(n_seq, 1, H, D)
    ks = k - kmean
    sq = q.abs().amax(dim=-1) / 127.0                                   # [n_seq, S, H]
    sq = torch.where(sq > 0, sq, torch.ones_like(sq))
    # code = rint(v * inv), inv = fp32(1 / s): the EXACT product (fp64 holds 24 x 24 bits) rounded once, half to even — what
    # the kernel's fused multiply-add with the 1.5 * 2^23 constant does; |v| <= the scale's maximum, so no clamp is needed
    code = lambda v, s: torch.round(v.double() * (1.0 / s).double()).float()
    q8 = code(q, sq[..., None])
    kb = ks.reshape(n_seq, S // BLOCK_K, BLOCK_K, H, D)
    sk = kb.abs().amax(dim=(2, 4)) / 127.0                              # [n_seq, S / 64, H]
    sk = torch.where(sk > 0, sk, torch.ones_like(sk))
    k8 = code(kb, sk[:, :, None, :, None]).reshape(n_seq, S, H, D)
    sv = v.abs().amax(dim=1) / 127.0                                    # [n_seq, H, D]
    sv = torch.where(sv > 0, sv, torch.ones_like(sv))
    v8 = code(v, sv[:, None])
    return dict(q8=q8, k8=k8, v8=v8, sq=sq, sk=sk, sv=sv, kmean=kmean[:, 0])


def attention_i8(qkv, n_seq, S, H, scale, D=72):
    """The scheme end to end; returns [n_seq * S, H * D] fp32."""
    z = quantise_qkv(qkv, n_seq, S, H, D)
    q8, k8, v8 = (z[n].permute(0, 2, 1, 3).double() for n in ("q8", "k8", "v8"))        # [n_seq, H, S, D]
    s_int = q8 @ k8.transpose(-1, -2)                                                   # exact integers
    sq = z["sq"].permute(0, 2, 1)[..., None].double()                                   # [n_seq, H, S, 1]
    sk = z["sk"].permute(0, 2, 1).repeat_interleave(BLOCK_K, dim=-1)[:, :, None, :].double()
    s = s_int * sq * sk * scale
    m = s.amax(dim=-1, keepdim=True)
    e = 255.0 * torch.exp(s - m)
    p8 = torch.round(e)
    o_int = p8 @ v8
    o = o_int * z["sv"][:, :, None, :].double() / e.sum(dim=-1, keepdim=True)      # normaliser: the UN-rounded sum
    return o.permute(0, 2, 1, 3).reshape(n_seq * S, H * D).float()


def attention_fp(qkv, n_seq, S, H, scale, D=72):
    """What the reference computes (blocks.py:169-188), in fp64 on the fp16 inputs."""
    x = qkv.double().reshape(n_seq, S, 3, H, D).permute(2, 0, 3, 1, 4)
    p = torch.softmax(x[0] @ x[1].transpose(-1, -2) * scale, dim=-1)
    return (p @ x[2]).permute(0, 2, 1, 3).reshape(n_seq * S, H * D).float()


def attention_i8_tiled(z, scale):
    """The kernel's own order of operations on the codes / scales `z` (quantise_qkv's dict), fp32 where the kernel
    computes in fp32: pass A takes the row maximum tile by tile (integer maximum x the tile's fp32 scale), pass B forms
    x = fma(S, c, log2(255) - m), P8 = rint(2^x), the exact integer P8 V8 and the fp32 row sum of the un-rounded 2^x."""
    q8, k8, v8 = (z[n].permute(0, 2, 1, 3).contiguous().double() for n in ("q8", "k8", "v8"))     # [n_seq, H, S, D]
    n_seq, H, S, D = q8.shape
    f32 = torch.float32
    scale_log2e = (torch.tensor(scale, dtype=f32) * torch.tensor(1.4426950408889634, dtype=f32))
    c_row = z["sq"].permute(0, 2, 1).to(f32) * scale_log2e                                # [n_seq, H, S]
    sk = z["sk"].permute(0, 2, 1).to(f32)                                                 # [n_seq, H, S / 64]
    lg255 = torch.tensor(7.994353436858858, dtype=f32)
    s_int = q8 @ k8.transpose(-1, -2)                                                     # exact integers [.., S, S]
    c_rt = (c_row[..., None] * sk[:, :, None, :])                                         # fp32 [n_seq, H, S, S / 64]
    tile_max = s_int.reshape(n_seq, H, S, S // BLOCK_K, BLOCK_K).amax(dim=-1).to(f32)
    m = (tile_max * c_rt).amax(dim=-1)                                                    # fp32 products, then the maximum
    neg = (lg255 - m)[..., None]
    x = (s_int * c_rt.repeat_interleave(BLOCK_K, dim=-1).double() + neg.double()).to(f32)  # one rounding, as the FMA
    e = torch.exp2(x.double()).to(f32)
    p8 = torch.round(e.double())                                                          # half to even, as the magic add
    l = e.double().sum(dim=-1, keepdim=True)
    out = (p8 @ v8) * z["sv"][:, :, None, :].double() / l
    return out.permute(0, 2, 1, 3).reshape(n_seq * S, H * D).float()
