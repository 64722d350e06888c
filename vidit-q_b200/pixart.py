"""PixArt-alpha / PixArtMS (t2i) graph around the same quantised linears — the second model behind the operator
(reference t2i/diffusion/model/nets/PixArtMS.py:47-211, PixArt_blocks.py:93-160).

state_dict names mirror the reference (x_embedder.proj, t_embedder.mlp.*, t_block.1, y_embedder.y_proj.fc{1,2},
blocks.N.{attn.{qkv,proj}, cross_attn.{q_linear,kv_linear,proj}, mlp.fc{1,2}, scale_shift_table}, final_layer.*), so
`viditq_b200.qdiff.QuantModel(model, ..., model_type="pixart")` picks QuantAttnLinearImg / QuantCrossAttnLinearImg by the
reference's rules.  CFG is a single batch-2n forward (dpm_solver model_fn), so per-token statistics pool over the
cond/uncond pair (quirk Q1) — the kernels take that as the pool group G.
`forward` = reference schedule (one QuantLayer call per linear); `forward_fused` = LN+modulate+quantise fused, GELU and
gated residuals in the GEMM epilogues, in-place residual stream.
"""
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .stdit import (CaptionEmbedder, Mlp, MultiHeadCrossAttention, STDiT, T2IFinalLayer, TimestepEmbedder,
                    _sincos_1d)

_ATTN_INT8_ENV = os.environ.get("VQ_ATTN_INT8", "0") == "1"


def pixart_pos_embed(dim, gh, gw, pe_interpolation=1.0, base_size=16):
    """PixArt.py:258-286."""
    h = np.arange(gh, dtype=np.float32) / (gh / base_size) / pe_interpolation
    w = np.arange(gw, dtype=np.float32) / (gw / base_size) / pe_interpolation
    grid = np.stack(np.meshgrid(w, h), axis=0).reshape(2, 1, gw, gh)
    return np.concatenate([_sincos_1d(dim // 2, grid[0]), _sincos_1d(dim // 2, grid[1])], axis=1)


class PatchEmbed(nn.Module):
    def __init__(self, patch_size, in_chans, embed_dim):
        super().__init__()
        self.patch_size = (patch_size, patch_size)
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class AttentionImg(nn.Module):
    """AttentionKVCompress without kv compression / qk-norm (PixArt-alpha defaults): fused qkv linear."""

    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)

    @staticmethod
    def attend(qkv, B, N, H, D):
        q5 = qkv.view(B, N, 3, H, D)
        o = F.scaled_dot_product_attention(q5[:, :, 0].transpose(1, 2), q5[:, :, 1].transpose(1, 2),
                                           q5[:, :, 2].transpose(1, 2))
        return o.transpose(1, 2).reshape(B, N, H * D)

    def forward(self, x):
        B, N, C = x.shape
        return self.proj(self.attend(self.qkv(x), B, N, self.num_heads, self.head_dim))


class PixArtMSBlock(nn.Module):
    def __init__(self, hidden_size, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.attn = AttentionImg(hidden_size, num_heads)
        self.cross_attn = MultiHeadCrossAttention(hidden_size, num_heads)
        self.norm2 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.mlp = Mlp(hidden_size, int(hidden_size * mlp_ratio))
        self.scale_shift_table = nn.Parameter(torch.randn(6, hidden_size) / hidden_size ** 0.5)

    def modulation(self, t):
        return (self.scale_shift_table[None] + t.reshape(t.shape[0], 6, -1)).chunk(6, dim=1)

    def forward(self, x, y, t, y_lens):
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = self.modulation(t)
        x = x + gate_msa * self.attn(self.norm1(x) * (1 + scale_msa) + shift_msa)
        x = x + self.cross_attn(x, y, y_lens)
        x = x + gate_mlp * self.mlp(self.norm2(x) * (1 + scale_mlp) + shift_mlp)
        return x


class PixArtMS(nn.Module):
    def __init__(self, input_size=64, patch_size=2, in_channels=4, hidden_size=1152, depth=28, num_heads=16,
                 mlp_ratio=4.0, pred_sigma=True, caption_channels=4096, pe_interpolation=1.0, model_max_length=120,
                 dtype=torch.float32, **unused):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = in_channels * 2 if pred_sigma else in_channels
        self.patch_size, self.hidden_size, self.num_heads, self.depth = patch_size, hidden_size, num_heads, depth
        self.pe_interpolation, self.base_size, self.dtype = pe_interpolation, input_size // patch_size, dtype
        n = (input_size // patch_size) ** 2
        self.register_buffer("pos_embed", torch.zeros(1, n, hidden_size))
        self.x_embedder = PatchEmbed(patch_size, in_channels, hidden_size)
        self.t_embedder = TimestepEmbedder(hidden_size)
        self.t_block = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 6 * hidden_size))
        self.y_embedder = CaptionEmbedder(caption_channels, hidden_size, model_max_length)
        self.blocks = nn.ModuleList([PixArtMSBlock(hidden_size, num_heads, mlp_ratio) for _ in range(depth)])
        self.final_layer = T2IFinalLayer(hidden_size, patch_size * patch_size, self.out_channels)
        self.init_synthetic()

    init_synthetic = STDiT.init_synthetic        # same seeded synthetic initialisation (all parameters covered)
    mask_select_plan = staticmethod(STDiT.mask_select_plan)
    kv_segments = staticmethod(STDiT.kv_segments)

    def _pos_embed(self, gh, gw, device):
        """PixArtMS recomputes its sin-cos table on the host every forward (PixArtMS.py:150-156); it only depends on the
        grid, so it is built once per (grid, device, dtype) and kept on the device."""
        key = (gh, gw, str(device), self.dtype)
        cache = self.__dict__.setdefault("_pe_cache", {})
        if key not in cache:
            pe = torch.from_numpy(pixart_pos_embed(self.hidden_size, gh, gw, self.pe_interpolation, self.base_size))
            cache[key] = pe.to(device).to(self.dtype).contiguous()
        return cache[key]

    def embed(self, x, timestep, y, mask, plan=None, fused=False):
        timestep = timestep.to(self.dtype)
        y = y.to(self.dtype)
        gh, gw = x.shape[-2] // self.patch_size, x.shape[-1] // self.patch_size
        pe = self._pos_embed(gh, gw, x.device)
        proj = self.x_embedder.proj
        if fused and proj.weight.dtype == torch.float16:
            # patchify + bias + position embedding in one pass (vq_patch_embed, the T = 1 case)
            x = ops.patch_embed(x.float().contiguous(), proj.weight, proj.bias, pe, (self.patch_size, self.patch_size))
        else:
            x = self.x_embedder(x.to(self.dtype)) + pe.unsqueeze(0)
        t = self.t_embedder(timestep, dtype=x.dtype)
        t0 = self.t_block(t)
        y = self.y_embedder(y)
        C = self.hidden_size
        if plan is not None:
            y_index, y_lens = plan
            y = y.squeeze(1).reshape(-1, C).index_select(0, y_index).view(1, -1, C)
        elif mask is not None:
            if mask.shape[0] != y.shape[0]:
                mask = mask.repeat(y.shape[0] // mask.shape[0], 1)
            mask = mask.squeeze(1).squeeze(1)
            y = y.squeeze(1).masked_select(mask.unsqueeze(-1) != 0).view(1, -1, C)
            y_lens = mask.sum(dim=1).tolist()
        else:
            y_lens = [y.shape[2]] * y.shape[0]
            y = y.squeeze(1).reshape(1, -1, C)
        return x, t, t0, y, y_lens

    def unpatchify(self, x):
        c, p = self.out_channels, self.patch_size
        h = w = int(x.shape[1] ** 0.5)
        x = x.reshape(x.shape[0], h, w, p, p, c)
        return torch.einsum("nhwpqc->nchpwq", x).reshape(x.shape[0], c, h * p, w * p)

    def forward(self, x, timestep, y, mask=None, data_info=None, **kwargs):
        x, t, t0, y, y_lens = self.embed(x, timestep, y, mask)
        for block in self.blocks:
            x = block(x, y, t0, y_lens)
        return self.unpatchify(self.final_layer(x, t))

    def forward_with_dpmsolver(self, x, timestep, y, data_info=None, **kwargs):
        return self.forward(x, timestep, y, data_info=data_info, **kwargs).chunk(2, dim=1)[0]

    def check_fused_state(self):
        """The fused schedule hard-wires dynamic per-token W+A quantisation of the 7 block linears (w8a8.yaml); LN-fused
        quantisers (qkv, fc1) cannot take smooth-quant scales.  Checked on every call; anything else -> forward()."""
        from .qdiff import QuantLayer, _is_dynamic
        for i, blk in enumerate(self.blocks):
            for path in ("attn.qkv", "attn.proj", "cross_attn.q_linear", "cross_attn.kv_linear", "cross_attn.proj",
                         "mlp.fc1", "mlp.fc2"):
                l = blk.get_submodule(path)
                if not (isinstance(l, QuantLayer) and l.weight_quant and l.act_quant and not l.disable_act_quant
                        and _is_dynamic(l.act_quantizer) and l.act_quantizer.per_group == "token"):
                    raise NotImplementedError(f"forward_fused: blocks.{i}.{path} is not dynamic per-token W+A quantised; "
                                              "use forward()")
                if path in ("attn.qkv", "mlp.fc1") and l.smooth_mode() is not None:
                    raise NotImplementedError(f"forward_fused: smooth-quant on the LN-fused blocks.{i}.{path}; use forward()")

    @staticmethod
    def _linear(layer, x, **kw):
        """One quantised linear of the fused schedule through vq_linear_w8a8 (pooled per-token statistics, quirk Q1)."""
        pw = layer._weight_for(x)
        x3 = x if x.dim() == 3 else x.reshape(1, -1, x.shape[-1])
        return ops.linear_w8a8(x3.contiguous(), pw, n_bits=layer.act_quantizer.n_bits, smooth=getattr(pw, "smooth", None), **kw)

    def forward_fused(self, x, timestep, y, mask=None, plan=None, segments=None):
        x, t, t0, y, y_lens = self.embed(x, timestep, y, mask, plan, fused=True)
        if segments is None:
            segments = self.kv_segments(y_lens, x.device)
        B, N, C = x.shape
        H, D = self.num_heads, C // self.num_heads
        M = B * N
        x = x.contiguous()
        xr = x.view(M, C)
        ones = torch.ones(1, C, dtype=x.dtype, device=x.device)
        self.check_fused_state()
        for blk in self.blocks:
            shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = (
                v.reshape(B, C).contiguous() for v in blk.modulation(t0))
            nb = blk.attn.qkv.act_quantizer.n_bits
            # every K = 1152 linear is ONE call (vq_linear_w8a8): at M = 2048 a single fused kernel each — LayerNorm +
            # modulate + quantise in the producer warps, dequant + bias / gated residual in the epilogue
            qkv = ops.linear_w8a8(x, blk.attn.qkv.prepared_weight(), n_bits=nb, ln=(shift_msa, scale_msa))
            if ops.attn_spatial_supported(N, D):   # tcgen05 flash attention, q|k|v read in place (one sequence per image)
                # VQ_ATTN_INT8=1 / .attn_int8: the opt-in INT8 Q/K/V attention (vq_attn_spatial_i8; own tolerance)
                attend = ops.attn_spatial_i8 if getattr(self, "attn_int8", _ATTN_INT8_ENV) else ops.attn_spatial
                o = attend(qkv, B, N, H, D, D ** -0.5).view(B, N, C)
            else:
                o = AttentionImg.attend(qkv, B, N, H, D)
            self._linear(blk.attn.proj, o, epi=ops.VQ_EPI_GATE_RESIDUAL, res=xr, gate=gate_msa, rows_per_gate=N, out=xr)
            ca = blk.cross_attn
            q = self._linear(ca.q_linear, x)
            kv = self._linear(ca.kv_linear, y)
            if D == 72 and max(y_lens) <= 128:
                o = ops.attn_cross(q, kv, segments[0], segments[1], B, N, H, D, max(y_lens), D ** -0.5).view(B, N, C)
            else:
                o = MultiHeadCrossAttention.attend(q, kv, B, N, y_lens, H, D).view(B, N, C)
            self._linear(ca.proj, o, epi=ops.VQ_EPI_GATE_RESIDUAL, res=xr, gate=ones, rows_per_gate=M, out=xr)
            h = ops.linear_w8a8(x, blk.mlp.fc1.prepared_weight(), n_bits=nb, ln=(shift_mlp, scale_mlp)).view(B, N, -1)
            # GELU rides in fc2's quantise pass (K = 4608: quantise pass + GEMM).  fc2 goes through its own
            # quantize_input: the running-stat smooth-quant EMA of quirk Q17 (PixArt keeps it on for blocks.27.mlp.fc2
            # at inference) sees gelu(h) there and re-quantises the weight for this call
            a = blk.mlp.fc2.quantize_input(h, gelu=True)
            ops.gemm_w8a8(a, a.pw, epi=ops.VQ_EPI_GATE_RESIDUAL, res=xr, gate=gate_mlp, rows_per_gate=N, out=xr)
        return self.unpatchify(self.final_layer(x, t))


def PixArtMS_XL_2(**kwargs):
    return PixArtMS(depth=28, hidden_size=1152, patch_size=2, num_heads=16, **kwargs)
