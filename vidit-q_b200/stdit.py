"""STDiT-XL/2 (OpenSORA v1.0 video DiT) graph around the quantised linears — the caller side of the hot path.

Parameter / buffer names mirror the reference (t2v/opensora/models/stdit/stdit.py:136-341, layers/blocks.py) so a
reference state_dict and a PTQ `ckpt.pth` load unchanged, and `viditq_b200.qdiff.QuantModel` applies the same
name-based layer replacement rules (quant_model.py:78-97).

Two forward schedules over the same parameters:
  * `forward`        — the reference's graph, one QuantLayer call per linear (13 per block, stdit.py:96-133).
  * `forward_fused`  — the B200 schedule: one LayerNorm+modulate+quantise pass feeds a single N=3456 q|k|v GEMM
                       (the reference quantises the same tensor three times, blocks.py:155-157), gated residuals and
                       GELU run in the GEMM epilogues, the temporal branch reads the (T S) token layout in place.
    Every rounding point of the reference's fp16 graph is kept, so both schedules agree to the last bit except for
    LayerNorm statistics (fp32 here, library-dependent there).
All three attentions are own kernels: spatial = tcgen05 flash attention (vq_attn_spatial), temporal / cross = small-sequence
kernels (vq_attn_temporal / vq_attn_cross); torch SDPA is only the library yardstick (VQ_SPATIAL_ATTN=sdpa).
"""
import math
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .qdiff import QuantLayer


def _sincos_1d(dim, pos):
    omega = np.arange(dim // 2, dtype=np.float64) / (dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", np.asarray(pos, dtype=np.float64).reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def _sincos_2d(dim, gh, gw, scale=1.0):
    h = np.arange(gh, dtype=np.float32) / scale
    w = np.arange(gw, dtype=np.float32) / scale
    grid = np.stack(np.meshgrid(w, h), axis=0).reshape(2, 1, gw, gh)  # w first, as in blocks.py:565-569
    return np.concatenate([_sincos_1d(dim // 2, grid[0]), _sincos_1d(dim // 2, grid[1])], axis=1)


class PatchEmbed3D(nn.Module):
    def __init__(self, patch_size, in_chans, embed_dim):
        super().__init__()
        self.patch_size = patch_size
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class TimestepEmbedder(nn.Module):
    def __init__(self, hidden_size, frequency_embedding_size=256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size), nn.SiLU(),
                                 nn.Linear(hidden_size, hidden_size))
        self.frequency_embedding_size = frequency_embedding_size

    def forward(self, t, dtype):
        half = self.frequency_embedding_size // 2
        freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
        args = t[:, None].float() * freqs[None]
        emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1).to(dtype)
        return self.mlp(emb)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features, out_features=None):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU(approximate="tanh")
        self.fc2 = nn.Linear(hidden_features, out_features or in_features)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class CaptionEmbedder(nn.Module):
    def __init__(self, in_channels, hidden_size, token_num=120):
        super().__init__()
        self.y_proj = Mlp(in_channels, hidden_size, hidden_size)
        self.register_buffer("y_embedding", torch.randn(token_num, in_channels) / in_channels ** 0.5)

    def forward(self, caption):
        return self.y_proj(caption)


class Attention(nn.Module):
    """Self-attention with separate q/k/v linears (blocks.py:113-195, separate_qkv=True)."""

    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.q = nn.Linear(dim, dim)
        self.k = nn.Linear(dim, dim)
        self.v = nn.Linear(dim, dim)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        shp = (B, N, self.num_heads, self.head_dim)
        q, k, v = (t.view(shp).transpose(1, 2) for t in (self.q(x), self.k(x), self.v(x)))
        o = F.scaled_dot_product_attention(q, k, v, scale=self.scale)
        return self.proj(o.transpose(1, 2).reshape(B, N, C))


class MultiHeadCrossAttention(nn.Module):
    """blocks.py:277-310: block-diagonal attention of image tokens over each sample's (mask-selected) prompt tokens."""

    def __init__(self, d_model, num_heads):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, d_model // num_heads
        self.q_linear = nn.Linear(d_model, d_model)
        self.kv_linear = nn.Linear(d_model, d_model * 2)
        self.proj = nn.Linear(d_model, d_model)

    @staticmethod
    def attend(q, kv, B, N, y_lens, H, D):
        """q [B*N, C]; kv [sum(y_lens), 2C] -> [B*N, C]."""
        q = q.view(B, N, H, D)
        kv = kv.view(-1, 2, H, D)
        outs, off = [], 0
        for b in range(B):
            L = y_lens[b]
            k = kv[off:off + L, 0].transpose(0, 1).unsqueeze(0)
            v = kv[off:off + L, 1].transpose(0, 1).unsqueeze(0)
            o = F.scaled_dot_product_attention(q[b].transpose(0, 1).unsqueeze(0), k, v)
            outs.append(o.squeeze(0).transpose(0, 1))
            off += L
        return torch.stack(outs, 0).reshape(B * N, H * D)

    def forward(self, x, cond, y_lens):
        B, N, C = x.shape
        q = self.q_linear(x).reshape(B * N, C)
        kv = self.kv_linear(cond).reshape(-1, 2 * C)
        o = self.attend(q, kv, B, N, y_lens, self.num_heads, self.head_dim)
        return self.proj(o.view(B, N, C))


class T2IFinalLayer(nn.Module):
    def __init__(self, hidden_size, num_patch, out_channels):
        super().__init__()
        self.norm_final = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.linear = nn.Linear(hidden_size, num_patch * out_channels)
        self.scale_shift_table = nn.Parameter(torch.randn(2, hidden_size) / hidden_size ** 0.5)
        self.out_channels = out_channels

    def forward(self, x, t):
        shift, scale = (self.scale_shift_table[None] + t[:, None]).chunk(2, dim=1)
        return self.linear(self.norm_final(x) * (1 + scale) + shift)


class STDiTBlock(nn.Module):
    def __init__(self, hidden_size, num_heads, d_s, d_t, mlp_ratio=4.0):
        super().__init__()
        self.hidden_size, self.d_s, self.d_t = hidden_size, d_s, d_t
        self.norm1 = nn.LayerNorm(hidden_size, eps=1e-6, elementwise_affine=False)
        self.attn = Attention(hidden_size, num_heads)
        self.cross_attn = MultiHeadCrossAttention(hidden_size, num_heads)
        self.norm2 = nn.LayerNorm(hidden_size, eps=1e-6, elementwise_affine=False)
        self.mlp = Mlp(hidden_size, int(hidden_size * mlp_ratio))
        self.scale_shift_table = nn.Parameter(torch.randn(6, hidden_size) / hidden_size ** 0.5)
        self.attn_temp = Attention(hidden_size, num_heads)

    def modulation(self, t):
        B = t.shape[0]
        return (self.scale_shift_table[None] + t.reshape(B, 6, -1)).chunk(6, dim=1)

    def forward(self, x, y, t, y_lens, tpe=None):
        """Reference schedule (stdit.py:96-133)."""
        B, N, C = x.shape
        T, S = self.d_t, self.d_s
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = self.modulation(t)
        x_m = self.norm1(x) * (1 + scale_msa) + shift_msa
        x_s = self.attn(x_m.view(B * T, S, C)).view(B, N, C)
        x = x + gate_msa * x_s
        x_t = x.view(B, T, S, C).transpose(1, 2).reshape(B * S, T, C)
        if tpe is not None:
            x_t = x_t + tpe
        x_t = self.attn_temp(x_t).view(B, S, T, C).transpose(1, 2).reshape(B, N, C)
        x = x + gate_msa * x_t
        x = x + self.cross_attn(x, y, y_lens)
        x = x + gate_mlp * self.mlp(self.norm2(x) * (1 + scale_mlp) + shift_mlp)
        return x


class STDiT(nn.Module):
    def __init__(self, input_size=(16, 64, 64), in_channels=4, patch_size=(1, 2, 2), hidden_size=1152, depth=28,
                 num_heads=16, mlp_ratio=4.0, pred_sigma=True, caption_channels=4096, model_max_length=120,
                 dtype=torch.float32, space_scale=1.0, time_scale=1.0, **unused):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = in_channels * 2 if pred_sigma else in_channels
        self.hidden_size, self.patch_size, self.input_size = hidden_size, patch_size, input_size
        self.num_temporal = input_size[0] // patch_size[0]
        self.num_patches = int(np.prod([input_size[i] // patch_size[i] for i in range(3)]))
        self.num_spatial = self.num_patches // self.num_temporal
        self.num_heads, self.depth, self.dtype = num_heads, depth, dtype
        gh, gw = input_size[1] // patch_size[1], input_size[2] // patch_size[2]
        self.register_buffer("pos_embed", torch.from_numpy(_sincos_2d(hidden_size, gh, gw, space_scale)).float()[None])
        tp = _sincos_1d(hidden_size, np.arange(self.num_temporal)[..., None] / time_scale)
        self.register_buffer("pos_embed_temporal", torch.from_numpy(tp).float()[None])
        self.x_embedder = PatchEmbed3D(patch_size, in_channels, hidden_size)
        self.t_embedder = TimestepEmbedder(hidden_size)
        self.t_block = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 6 * hidden_size))
        self.y_embedder = CaptionEmbedder(caption_channels, hidden_size, model_max_length)
        self.blocks = nn.ModuleList([STDiTBlock(hidden_size, num_heads, self.num_spatial, self.num_temporal, mlp_ratio)
                                     for _ in range(depth)])
        self.final_layer = T2IFinalLayer(hidden_size, int(np.prod(patch_size)), self.out_channels)
        self.init_synthetic()

    @torch.no_grad()
    def init_synthetic(self, seed=0):
        """Seeded synthetic weights (no checkpoints in this environment): xavier linears, and — unlike the reference's
        zero-init of cross_attn.proj / attn_temp.proj / final_layer.linear (stdit.py:409-452) — non-zero everywhere so
        that every branch contributes to the output in parity tests."""
        g = torch.Generator().manual_seed(seed)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                bound = math.sqrt(6.0 / (m.in_features + m.out_features))
                m.weight.copy_((torch.rand(m.weight.shape, generator=g) * 2 - 1) * bound)
                if m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.02)
        w = self.x_embedder.proj.weight
        bound = math.sqrt(6.0 / (w[0].numel() + w.shape[0]))
        w.copy_((torch.rand(w.shape, generator=g) * 2 - 1) * bound)
        for b in self.blocks:
            b.scale_shift_table.copy_(torch.randn(b.scale_shift_table.shape, generator=g) / self.hidden_size ** 0.5)
        fl = self.final_layer.scale_shift_table
        fl.copy_(torch.randn(fl.shape, generator=g) / self.hidden_size ** 0.5)
        ye = self.y_embedder.y_embedding
        ye.copy_(torch.randn(ye.shape, generator=g) / ye.shape[1] ** 0.5)
        self.x_embedder.proj.bias.copy_(torch.randn(self.x_embedder.proj.bias.shape, generator=g) * 0.02)

    # ---- shared pre/post --------------------------------------------------------------------------------------------
    @staticmethod
    def mask_select_plan(mask):
        """Host-side plan of the MASK_SELECT gather (stdit.py:280-286) for a fixed prompt mask [B, L]: flat indices of
        the kept tokens and the per-sample lengths. Computed once per prompt so the forward itself is sync-free
        (CUDA-graph capturable)."""
        m = mask.reshape(mask.shape[0], -1).cpu()
        idx = torch.nonzero(m.reshape(-1) != 0).reshape(-1)
        return idx.to(mask.device), [int(v) for v in m.sum(dim=1).tolist()]

    @staticmethod
    def kv_segments(y_lens, device):
        """Device int32 (start, length) of every sample's prompt rows in the packed kv tensor."""
        starts = [0]
        for v in y_lens[:-1]:
            starts.append(starts[-1] + v)
        return (torch.tensor(starts, dtype=torch.int32, device=device),
                torch.tensor(list(y_lens), dtype=torch.int32, device=device))

    def embed(self, x, timestep, y, mask, plan=None, fused=False):
        timestep = timestep.to(self.dtype)
        y = y.to(self.dtype)
        B = x.shape[0]
        T, S, C = x.shape[2] // self.patch_size[0], self.num_spatial, self.hidden_size   # T: this rank's frames
        proj = self.x_embedder.proj
        if fused and self.patch_size[0] == 1 and proj.weight.dtype == torch.float16:
            # one pass: patchify + bias + "B (T S) C" layout + pos_embed (vq_patch_embed) instead of conv3d, two layout
            # conversions, a transposing add
            x = ops.patch_embed(x.float().contiguous(), proj.weight, proj.bias, self.pos_embed.to(torch.float16).reshape(S, C),
                                self.patch_size[1:])
        else:
            x = x.to(self.dtype)
            x = self.x_embedder(x).view(B, T, S, C) + self.pos_embed
            x = x.view(B, T * S, C)
        t = self.t_embedder(timestep, dtype=x.dtype)
        t0 = self.t_block(t)
        y = self.y_embedder(y)
        aq = getattr(self.final_layer.linear, "act_quantizer", None)
        mask_select = not (aq is not None and type(aq).__name__ != "DynamicActQuantizer" and aq.per_group == "token")
        if not mask_select and (fused or plan is not None):
            # static per-token scales are calibrated per position of each layer's own (pooled) view — the temporal
            # layers' in (S T) order — which the fused schedule's in-place (T S) layout does not reproduce
            raise NotImplementedError("static per-token activation scales (MASK_SELECT=False, quirk Q14): use forward()")
        if plan is not None:
            y_index, y_lens = plan
            y = y.squeeze(1).reshape(-1, C).index_select(0, y_index).view(1, -1, C)
        elif mask is not None and mask_select:  # MASK_SELECT branch of stdit.py:280-286 (dynamic activation quantiser)
            if mask.shape[0] != y.shape[0]:
                mask = mask.repeat(y.shape[0] // mask.shape[0], 1)
            mask = mask.squeeze(1).squeeze(1)
            y = y.squeeze(1).masked_select(mask.unsqueeze(-1) != 0).view(1, -1, C)
            y_lens = mask.sum(dim=1).tolist()
        elif mask is not None:   # MASK_SELECT = False (stdit.py:287-300, quirk Q14): padded prompt rows are zeroed, not
            # dropped, so every static per-token (delta, zp) keeps its position: always max_len rows per sample
            mask_ = mask.repeat(y.shape[0] // mask.shape[0], 1) if mask.shape[0] != y.shape[0] else mask
            y_lens = [y.shape[2]] * y.shape[0]
            y = (y * mask_.unsqueeze(-1).unsqueeze(1).to(y.dtype)).squeeze(1).reshape(1, -1, C)
        else:
            y_lens = [y.shape[2]] * y.shape[0]
            y = y.squeeze(1).reshape(1, -1, C)
        return x, t, t0, y, y_lens

    def unpatchify(self, x):
        B = x.shape[0]
        Nt, Nh, Nw = [self.input_size[i] // self.patch_size[i] for i in range(3)]
        Tp, Hp, Wp = self.patch_size
        Nt = x.shape[1] // (Nh * Nw)      # frame-sharded forward: only this rank's frames
        x = x.view(B, Nt, Nh, Nw, Tp, Hp, Wp, self.out_channels)
        return x.permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(B, self.out_channels, Nt * Tp, Nh * Hp, Nw * Wp)

    def forward(self, x, timestep, y, mask=None, plan=None):
        """The reference's schedule, layer by layer (stdit.py:246-312).  plan: a host-precomputed mask_select_plan(mask)
        replaces the masked_select + `.tolist()` of the MASK_SELECT branch, which makes the call free of host syncs — the
        whole layer-by-layer forward can then be captured in a CUDA graph (bench.py --schedule hook-graph)."""
        x, t, t0, y, y_lens = self.embed(x, timestep, y, mask, plan)
        tpe = self.pos_embed_temporal.to(x.dtype)
        for i, block in enumerate(self.blocks):
            x = block(x, y, t0, y_lens, tpe if i == 0 else None)
        x = self.final_layer(x, t)
        return self.unpatchify(x).to(torch.float32)

    # ---- fused B200 schedule ----------------------------------------------------------------------------------------
    def forward_fused(self, x, timestep, y, mask=None, plan=None, segments=None, independent=False, frames=None):
        """plan / segments: host-precomputed mask_select_plan(mask) and kv_segments(y_lens) make the call sync-free.
        independent=True: the batch entries are SEPARATE reference forward calls stacked into one launch sequence — the
        cond / uncond halves of cfg_split=True (iddpm/__init__.py:156-157 calls the model twice with batch n_prompts = 1).
        Their per-token statistics are not pooled (each row is quantised on its own, which is what two batch-1 calls do);
        results are identical to calling forward_fused once per entry, at half the launches and better-filled GEMM waves.
        frames = (process group, P, rank in group): frame-sharded forward of one video — x holds this rank's T / P frames
        of the latent, the returned tensor the same frames of the output (viditq_b200.shard)."""
        x, t, t0, y, y_lens = self.embed(x, timestep, y, mask, plan, fused=True)
        eng = getattr(self, "_engine", None)
        if eng is None:
            eng = self._engine = FusedBlocks(self)
        if segments is None:
            segments = self.kv_segments(y_lens, x.device)
        x = eng.run(x, y, t0, y_lens, segments, independent, frames)
        # final layer (FP, remain_fp.txt): LayerNorm + modulate in one pass of the fused kernel (its codes are unused)
        fl = self.final_layer
        shift, scale = (fl.scale_shift_table[None] + t[:, None]).chunk(2, dim=1)
        B, _, C = x.shape
        _, xm = ops.ln_modulate_act_quant(x, shift.reshape(B, C).contiguous(), scale.reshape(B, C).contiguous(), want_y=True)
        x = fl.linear(xm)
        return self.unpatchify(x).to(torch.float32)


def STDiT_XL_2(**kwargs):
    return STDiT(depth=28, hidden_size=1152, patch_size=(1, 2, 2), num_heads=16, **kwargs)


class FusedBlocks:
    """Block-level schedule on the fused kernels. Reads the QuantLayers' prepared weights; q|k|v weights of each
    attention are concatenated once into one [3C, K] operand (identical per-channel parameters, one GEMM)."""

    def __init__(self, model: STDiT):
        self.m = model
        self._qkv = {}
        # VQ_SPATIAL_ATTN=sdpa runs the long spatial attention on the library flash kernel (torch SDPA -> cuDNN) instead of
        # vq_attn_spatial: the yardstick bench.py / tools/prof_kernels.py time the own kernel against, not a fallback
        self.own_spatial = os.environ.get("VQ_SPATIAL_ATTN", "own") != "sdpa"
        # VQ_LINEAR_FUSED=2: every K = 1152 linear of the schedule goes through the ONE-call entry vq_linear_w8a8 in its
        # overlapped mode — the persistent GEMM with quantiser warpgroups running ahead of its MMAs (no separate quantise pass)
        self.overlap = os.environ.get("VQ_LINEAR_FUSED", "0") == "2"
        # VQ_TEMPORAL_FUSED_QUANT=1: temporal attention + the projection's quantiser as ONE kernel (vq_attn_temporal_quant).
        # Bit-identical; measured on B200: 101 us against 62 + 18 us for the two kernels in a single replay (the 16-warp,
        # 135 KB block runs one per SM: no second block to cover its load and barrier phases), equal step time under the
        # power cap — so the two-kernel sequence stays the default
        self.temporal_fused_quant = os.environ.get("VQ_TEMPORAL_FUSED_QUANT", "0") == "1"
        # VQ_ATTN_INT8=1 (or .attn_int8 = True): OPT-IN spatial attention that consumes INT8 Q/K/V on tcgen05 kind::i8
        # (vq_attn_spatial_i8).  The reference keeps attention in fp16 (its Q/K/V quantisers are commented out,
        # quant_block.py:617-632), so this leaves the reference's numerics: own tolerance, DESIGN.md 4.2d.  Default off.
        self.attn_int8 = os.environ.get("VQ_ATTN_INT8", "0") == "1"
        self.static = False          # set by check_state(): all block linears carry static per-tensor activation scales
        self._static_same = {}       # (block, branch) -> (key, q / k / v share one calibrated (delta, zero point))

    @staticmethod
    def _spatial_library(qkv, pj, scale, B, N, T, S, C, D, independent, static=False):
        """Library yardstick / shapes vq_attn_spatial does not cover (head_dim != 72 or S not a multiple of 256)."""
        o = F.scaled_dot_product_attention(qkv[:, :, 0].transpose(1, 2), qkv[:, :, 1].transpose(1, 2),
                                           qkv[:, :, 2].transpose(1, 2), scale=scale)   # [B*T, H, S, D]
        if o.is_contiguous() and D == 72 and C == 1152 and not pj.smooth_quant and not static:
            # quantise straight from the head-major layout the library kernel emits (no transpose copy)
            a = (ops.act_quant_heads(o, 1, B * N, S, n_bits=pj.act_quantizer.n_bits) if independent
                 else ops.act_quant_heads(o, B, N, S, n_bits=pj.act_quantizer.n_bits))
            a.pw = pj.prepared_weight()
            return a
        return pj.quantize_input(o.transpose(1, 2).reshape(B * T, S, C), independent=independent)

    LN_FUSED = ("attn.q", "attn.k", "attn.v", "attn_temp.q", "attn_temp.k", "attn_temp.v", "mlp.fc1")

    def check_state(self):
        """The fused schedule runs all 13 block linears W+A quantised, their activations either per-token DYNAMIC (the
        ViDiT-Q W8A8 / W4A8 configs: quantisers fused with LayerNorm / add / GELU passes) or — all of them — with STATIC
        per-tensor calibrated scales (w8a8_naive.yaml: the transformed fp16 tensor is formed first, then vq_act_quant_static).
        Anything else — a layer switched to FP by set_layer_quant, static per-token scales, a mix of the two kinds,
        input-dependent smooth-quant scales in front of an LN-fused quantiser, q/k/v with different activation widths —
        must go through STDiT.forward (one QuantLayer call per linear); checked on EVERY call.  Sets self.static."""
        from .qdiff import _is_dynamic
        kinds = set()
        for i, blk in enumerate(self.m.blocks):
            for path in ("attn.q", "attn.k", "attn.v", "attn.proj", "attn_temp.q", "attn_temp.k", "attn_temp.v",
                         "attn_temp.proj", "cross_attn.q_linear", "cross_attn.kv_linear", "cross_attn.proj", "mlp.fc1",
                         "mlp.fc2"):
                l = blk.get_submodule(path)
                if not (isinstance(l, QuantLayer) and l.weight_quant and l.act_quant and not l.disable_act_quant):
                    raise NotImplementedError(f"forward_fused: blocks.{i}.{path} is not in W+A quantised state; use forward()")
                aq = l.act_quantizer
                if _is_dynamic(aq) and aq.per_group == "token":
                    kinds.add("dynamic")
                elif (not _is_dynamic(aq) and not aq.per_group and aq.init_done and aq.delta is not None
                      and aq.delta.numel() == 1):
                    kinds.add("static")
                else:
                    raise NotImplementedError(f"forward_fused: blocks.{i}.{path} has static per-token / un-calibrated / "
                                              "non-per-token activation scales; use forward()")
                if len(kinds) > 1:
                    raise NotImplementedError(f"forward_fused: blocks.{i}.{path} mixes static and dynamic activation "
                                              "quantisers with the layers before it; use forward()")
                if path in self.LN_FUSED and l.smooth_mode() in ("dynamic", "running"):
                    raise NotImplementedError(f"forward_fused: blocks.{i}.{path} uses an input-dependent smooth-quant scale "
                                              "in front of a fused LayerNorm / add quantiser; use forward()")
            for att in (blk.attn, blk.attn_temp):
                if len({att.q.act_quantizer.n_bits, att.k.act_quantizer.n_bits, att.v.act_quantizer.n_bits}) != 1:
                    raise NotImplementedError(f"forward_fused: blocks.{i} q/k/v activation widths differ (shared quantise "
                                              "pass); use forward()")
        self.static = kinds == {"static"}

    def _qkv_weight(self, attn, tag):
        layers = (attn.q, attn.k, attn.v)
        if any(l.smooth_quant for l in layers):
            return None   # per-layer channel scales (quant_layer.py:137 depends on each weight): no shared input codes
        pws = [l.prepared_weight() for l in layers]   # validates each layer's cache (bumps _gen when its source moved)
        key = (tag,) + tuple((l.weight_quantizer.n_bits, l._gen) for l in layers)
        hit = self._qkv.get(tag)
        if hit is None or hit[0] != key:
            codes = torch.cat([p.codes for p in pws], 0).contiguous()
            col = torch.cat([p.col for p in pws], 0).contiguous()
            hit = self._qkv[tag] = (key, ops.PreparedWeight(codes, col, codes.shape[0], pws[0].K, pws[0].n_bits))
        return hit[1]

    def _static_qkv(self, attn, tag, t, ln=None, rpm=None):
        """Static per-tensor scales: `t` is the fp16 tensor in front of the three layers; ln = (shift, scale) has the static
        quantiser apply LayerNorm + modulate on the fly (vq_ln_modulate_act_quant_static).  q, k and v were calibrated on
        the same tensor; when their (delta, zero point) are identical and no per-layer smooth-quant scale is in the way, ONE
        quantise pass feeds the concatenated N = 3C GEMM, otherwise each layer quantises for itself and the three GEMMs
        write column slices of one output."""
        layers = (attn.q, attn.k, attn.v)
        C = t.shape[-1]
        pw = self._qkv_weight(attn, tag)
        if pw is not None:
            aqs = [l.act_quantizer for l in layers]
            key = tuple((q.delta.data_ptr(), q.delta._version, q.zero_point.data_ptr(), q.zero_point._version) for q in aqs)
            hit = self._static_same.get(tag)
            if hit is None or hit[0] != key:   # one host comparison per checkpoint (outside any graph capture: warm-up call)
                same = all(torch.equal(aqs[0].delta.reshape(-1), q.delta.reshape(-1)) and
                           torch.equal(aqs[0].zero_point.reshape(-1), q.zero_point.reshape(-1)) for q in aqs[1:])
                hit = self._static_same[tag] = (key, same)
            if hit[1]:
                return ops.gemm_w8a8(layers[0].quantize_input(t, ln=ln, rows_per_mod=rpm), pw)
        out = torch.empty(t.numel() // C, 3 * C, dtype=t.dtype, device=t.device)
        for j, layer in enumerate(layers):
            a = layer.quantize_input(t, ln=ln, rows_per_mod=rpm)
            ops.gemm_w8a8(a, a.pw, out=out[:, j * C:(j + 1) * C], ldo=3 * C)
        return out

    def _qlin(self, layer, t, qi, unpooled, **kw):
        """Quantiser + GEMM of one block linear on `t` [.., n, C].  Overlapped mode (un-pooled statistics only): one call,
        the quantise arithmetic runs inside the GEMM kernel."""
        if self.overlap and unpooled and not self.static and layer.smooth_mode() in (None, "cached"):
            pw = layer.prepared_weight()
            return ops.linear_w8a8(t.reshape(1, -1, t.shape[-1]), pw, n_bits=layer.act_quantizer.n_bits,
                                   smooth=getattr(pw, "smooth", None), **kw)
        a = qi(layer, t)
        return ops.gemm_w8a8(a, a.pw, **kw)

    def _qkv_project(self, attn, tag, x, ln=None, independent=False, add=None):
        """q|k|v of one attention as one [M, 3C] tensor. ln = (shift, scale) fuses LayerNorm+modulate in front.
        Without smooth-quant: one quantise pass + one N=3C GEMM. With it (w4a8_timestep_aware_cb.yaml): each layer has
        its own channel scale, hence its own codes; the three GEMMs write column slices of the same output.
        add = (vectors [period, C], rows_per_add): an fp16 row-broadcast add fused in front of the quantiser (block 0's
        temporal position embedding) when K = 1152."""
        if self.static:   # calibrated scales: LayerNorm + modulate inside the static quantiser; adds formed first
            if ln is not None:
                return self._static_qkv(attn, tag, x, ln=ln, rpm=x.shape[1])
            if add is not None:
                vec, rpa = add
                x = (x.reshape(-1, vec.shape[0], rpa, x.shape[-1]) + vec.view(1, -1, 1, x.shape[-1])).view(x.shape)
            return self._static_qkv(attn, tag, x)
        pw = self._qkv_weight(attn, tag)
        nb = attn.q.act_quantizer.n_bits
        rpm = None
        if independent:   # stacked separate calls: un-pooled statistics, per-entry modulation vectors
            rpm = x.shape[1]
            x = x.view(1, -1, x.shape[2])
        if pw is not None:
            if self.overlap and add is None and x.shape[0] == 1:
                return ops.linear_w8a8(x, pw, n_bits=nb, ln=ln, rows_per_mod=rpm if ln is not None else None)
            if ln is not None:
                a = ops.ln_modulate_act_quant(x, ln[0], ln[1], n_bits=nb, rows_per_mod=rpm)[0]
            elif add is not None:
                a = ops.add_act_quant(x, add[0], add[1], n_bits=nb)
            else:
                a = ops.act_quant(x, n_bits=nb)
            return ops.gemm_w8a8(a, pw)
        B, N, C = x.shape
        out = torch.empty(B * N, 3 * C, dtype=x.dtype, device=x.device)
        for j, layer in enumerate((attn.q, attn.k, attn.v)):
            lw = layer.prepared_weight()
            sm = getattr(lw, "smooth", None)
            if ln is not None:
                a = ops.ln_modulate_act_quant(x, ln[0], ln[1], n_bits=nb, smooth=sm, rows_per_mod=rpm)[0]
            elif add is not None:
                a = ops.add_act_quant(x, add[0], add[1], n_bits=nb, smooth=sm)
            else:
                a = ops.act_quant(x, n_bits=nb, smooth=sm)
            ops.gemm_w8a8(a, lw, out=out[:, j * C:(j + 1) * C], ldo=3 * C)
        return out

    def run(self, x, y, t0, y_lens, segments, independent=False, frames=None):
        m = self.m
        self.check_state()
        B, N, C = x.shape
        S, H = m.num_spatial, m.num_heads
        T = N // S                       # frames held by this rank (all of them unless frame-sharded)
        D = C // H
        M = B * N
        t_first = 0
        if frames is not None:
            from . import shard
            grp, P, prank = frames
            if not (independent or B == 1):
                raise NotImplementedError("frame sharding moves per-token codes: needs un-pooled statistics")
            if self.static:
                raise NotImplementedError("frame sharding exchanges per-token codes with their dynamic (delta, zp) tails")
            if T * P != m.num_temporal or S % P:
                raise ValueError(f"frame sharding: {T} local frames x {P} ranks != {m.num_temporal} or {S} % {P} != 0")
            t_first = prank * T

        def qi(layer, t, gelu=False):
            """The layer's activation quantiser. Frame-sharded: a rank holds T / P frames, which the reference layers'
            (B, T*S) pooling views do not describe — every row is quantised on its own (un-pooled, as checked above)."""
            if self.static:   # calibrated scales: GELU rides in the static quantise pass (vq_gelu_act_quant_static)
                return layer.quantize_input(t, gelu=gelu)
            if frames is None:
                return layer.quantize_input(t, gelu=gelu, independent=independent)
            if layer.smooth_quant:
                raise NotImplementedError("frame sharding with smooth-quant channel scales")
            a_ = ops.act_quant(t.reshape(1, -1, t.shape[-1]), n_bits=layer.act_quantizer.n_bits, gelu=gelu)
            a_.pw = layer.prepared_weight()
            return a_
        x = x.contiguous()   # fresh tensor from embed(): the residual stream is updated in place below
        ones = torch.ones(1, C, dtype=x.dtype, device=x.device)
        tpe = m.pos_embed_temporal.to(x.dtype)
        # all blocks' t2i modulation vectors in three launches: [L, 6, B, C] = scale_shift_table[None] + t0 (stdit.py:100-102)
        tables = torch.stack([blk.scale_shift_table for blk in m.blocks]).to(x.dtype)
        mod = (tables[:, None] + t0.reshape(1, B, 6, C)).permute(0, 2, 1, 3).contiguous()
        for i, blk in enumerate(m.blocks):
            shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = mod[i].unbind(0)
            # ---- spatial attention: LN + modulate + quantise once, one q|k|v GEMM
            qkv = self._qkv_project(blk.attn, (i, "s"), x, ln=(shift_msa, scale_msa), independent=independent)
            pj = blk.attn.proj
            unpooled = frames is None and (independent or B == 1)
            xr = x.view(M, C)   # residual stream, updated in place: out aliases res
            if self.own_spatial and ops.attn_spatial_supported(S, D):
                # tcgen05 flash attention reading q|k|v in place, token-major output: the projection's quantiser input
                o = (ops.attn_spatial_i8 if self.attn_int8 else ops.attn_spatial)(qkv, B * T, S, H, D, blk.attn.scale)
                self._qlin(pj, o.view(B * T, S, C), qi, unpooled, epi=ops.VQ_EPI_GATE_RESIDUAL, res=xr, gate=gate_msa,
                           rows_per_gate=N, out=xr)
            else:
                a = self._spatial_library(qkv.view(B * T, S, 3, H, D), pj, blk.attn.scale, B, N, T, S, C, D, independent,
                                          self.static)
                ops.gemm_w8a8(a, a.pw, epi=ops.VQ_EPI_GATE_RESIDUAL, res=xr, gate=gate_msa, rows_per_gate=N, out=xr)
            # ---- temporal attention on the (T S) layout (+ temporal pos-emb in block 0)
            if frames is not None:
                # frame-sharded: quantise locally, all-to-all the CODES into the (all frames, S / P positions) layout,
                # q|k|v GEMM + attention + the projection's quantiser there, all-to-all the codes back
                pw_t = self._qkv_weight(blk.attn_temp, (i, "t"))
                if pw_t is None or blk.attn_temp.proj.smooth_quant or C != 1152:
                    raise NotImplementedError("frame sharding with per-layer smooth-quant scales")
                nb = blk.attn_temp.q.act_quantizer.n_bits
                x1 = x.view(1, M, C)
                a = (ops.add_act_quant(x1, tpe.view(-1, C)[t_first:t_first + T].contiguous(), S, n_bits=nb) if i == 0
                     else ops.act_quant(x1, n_bits=nb))
                a = shard.exchange_act_codes(a, B, T, S, P, True, grp)
                qkv = ops.gemm_w8a8(a, pw_t)
                nbp = blk.attn_temp.proj.act_quantizer.n_bits
                if self.temporal_fused_quant and ops.attn_temporal_quant_supported(T * P, H, D):
                    a = ops.attn_temporal_quant(qkv, B, T * P, S // P, H, D, blk.attn_temp.scale, n_bits=nbp)
                else:
                    o = ops.attn_temporal(qkv, B, T * P, S // P, H, D, blk.attn_temp.scale)
                    a = ops.act_quant(o.view(1, -1, C), n_bits=nbp)
                a = shard.exchange_act_codes(a, B, T, S, P, False, grp)
                ops.gemm_w8a8(a, blk.attn_temp.proj.prepared_weight(), epi=ops.VQ_EPI_GATE_RESIDUAL, res=xr,
                              gate=gate_msa, rows_per_gate=N, out=xr)
            elif i == 0 and C == 1152 and not self.static:   # x + tpe rides in the quantise pass (vq_add_act_quant): frame t = (row // S) % T
                qkv = self._qkv_project(blk.attn_temp, (i, "t"), x, independent=independent, add=(tpe.view(T, C), S))
            else:
                xt = x if i != 0 else (x.view(B, T, S, C) + tpe.view(1, T, 1, C)).view(B, N, C)
                qkv = self._qkv_project(blk.attn_temp, (i, "t"), xt, independent=independent)
            pjt = blk.attn_temp.proj
            if (frames is None and self.temporal_fused_quant and (independent or B == 1) and not self.static
                    and ops.attn_temporal_quant_supported(T, H, D) and pjt.smooth_mode() in (None, "cached")):
                # temporal attention + the projection's quantiser in ONE kernel (all 16 heads of a position in one block):
                # no fp16 attention output, no separate quantise pass
                pwt = pjt.prepared_weight()
                a = ops.attn_temporal_quant(qkv, B, T, S, H, D, blk.attn_temp.scale, n_bits=pjt.act_quantizer.n_bits,
                                            smooth=getattr(pwt, "smooth", None))
                ops.gemm_w8a8(a, pwt, epi=ops.VQ_EPI_GATE_RESIDUAL, res=xr, gate=gate_msa, rows_per_gate=N, out=xr)
            elif frames is None:
                if T <= 16 and D == 72:      # own kernel: reads the (T S) layout in place, no permute copies
                    o = ops.attn_temporal(qkv, B, T, S, H, D, blk.attn_temp.scale).view(B, N, C)
                else:                        # library path for shapes the kernel does not cover
                    q5 = qkv.view(B, T, S, 3, H, D)
                    qt, kt, vt = (q5[:, :, :, j].permute(0, 2, 3, 1, 4).reshape(B * S, H, T, D) for j in range(3))
                    o = F.scaled_dot_product_attention(qt, kt, vt, scale=blk.attn_temp.scale)
                    o = o.view(B, S, H, T, D).permute(0, 3, 1, 2, 4).reshape(B, N, C)
                # per-token statistics: row order irrelevant
                self._qlin(blk.attn_temp.proj, o.view(B * S, T, C),
                           lambda l, t_: l.quantize_input(t_, independent=independent), unpooled,
                           epi=ops.VQ_EPI_GATE_RESIDUAL, res=xr, gate=gate_msa, rows_per_gate=N, out=xr)
            # ---- cross attention
            ca = blk.cross_attn
            q = self._qlin(ca.q_linear, x, qi, unpooled)
            a = ca.kv_linear.quantize_input(y)
            kv = ops.gemm_w8a8(a, a.pw)
            if D == 72 and max(y_lens) <= 128:
                o = ops.attn_cross(q, kv, segments[0], segments[1], B, N, H, D, max(y_lens), D ** -0.5).view(B, N, C)
            else:
                o = MultiHeadCrossAttention.attend(q, kv, B, N, y_lens, H, D).view(B, N, C)
            self._qlin(ca.proj, o, qi, unpooled, epi=ops.VQ_EPI_GATE_RESIDUAL, res=xr, gate=ones, rows_per_gate=M, out=xr)
            # ---- MLP: LN + modulate + quantise, fc1 (+GELU), quantise, fc2 (+gate, residual)
            fc1w = blk.mlp.fc1.prepared_weight()
            # GELU rides in fc2's quantise pass (HBM-bound, idle MUFU) instead of fc1's epilogue (epilogue-bound)
            if self.static:
                a = blk.mlp.fc1.quantize_input(x, ln=(shift_mlp, scale_mlp), rows_per_mod=N)
                h = ops.gemm_w8a8(a, a.pw).view(B, N, -1)
            elif self.overlap and unpooled:
                h = ops.linear_w8a8(x.view(1, M, C), fc1w, n_bits=blk.mlp.fc1.act_quantizer.n_bits,
                                    smooth=getattr(fc1w, "smooth", None), ln=(shift_mlp, scale_mlp), rows_per_mod=N).view(B, N, -1)
            else:
                a, _ = ops.ln_modulate_act_quant(x.view(1, M, C) if independent else x, shift_mlp, scale_mlp,
                                                 n_bits=blk.mlp.fc1.act_quantizer.n_bits,
                                                 smooth=getattr(fc1w, "smooth", None), rows_per_mod=N if independent else None)
                h = ops.gemm_w8a8(a, fc1w).view(B, N, -1)
            a = qi(blk.mlp.fc2, h, gelu=True)
            ops.gemm_w8a8(a, a.pw, epi=ops.VQ_EPI_GATE_RESIDUAL, res=xr, gate=gate_mlp, rows_per_gate=N, out=xr)
        return x
