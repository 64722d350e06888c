"""Import shims that let the UNMODIFIED reference (thu-nics/ViDiT-Q, /root/reference) be imported in this container.

TEST INFRASTRUCTURE ONLY (golden-vector generation and oracle validation; /root/reference does not exist on the GPU box,
so nothing on the `-m gpu` / smoke / bench paths calls this).  No reference source is copied: the shims only provide the
third-party modules the reference imports but this image lacks (omegaconf, diffusers, timm, xformers, mmengine registry,
and qdiff.models.quant_block, which needs diffusers 0.24 internals and is dead code for STDiT/PixArt — SURVEY.md §2 #7).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VIDITQ_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "qdiff"))


class AttrDict(dict):
    """Quantiser config: the reference reads it both as attributes and via .get() (base_quantizer.py:29-49)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Idempotently register the stubs and put the reference on sys.path."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "omegaconf" not in sys.modules:
        class ListConfig(list):
            pass
        _module("omegaconf", ListConfig=ListConfig)
    if "diffusers" not in sys.modules:
        _module("diffusers")
    if "qdiff.models.quant_block" not in sys.modules:
        import torch.nn as nn

        class BaseQuantBlock(nn.Module):
            pass

        class TransformerBlock(nn.Module):
            pass

        class QuantTransformerBlock(BaseQuantBlock):
            pass

        def get_specials(*a, **k):
            return []
        import importlib
        importlib.import_module("qdiff.models")  # namespace package from the reference tree
        _module("qdiff.models.quant_block", BaseQuantBlock=BaseQuantBlock, TransformerBlock=TransformerBlock,
                QuantTransformerBlock=QuantTransformerBlock, get_specials=get_specials)


def w8a8_dynamic_configs(n_temporal=16, n_spatial=1024, n_prompt=120, w_bits=8, a_bits=8, smooth=None):
    """The quantiser sections of t2v/configs/quant/opensora/w8a8_dynamic.yaml (and w4a8_timestep_aware_cb.yaml when
    `smooth` is given) as attribute dicts."""
    wq = AttrDict(n_bits=w_bits, channel_wise=True, per_group="channel", channel_dim=0, scale_method="min_max",
                  round_mode="nearest")
    sq = AttrDict(enable=False, channel_wise_scale_type="momentum_act_max", momentum=0.95, alpha=0.625)
    if smooth is not None:
        sq = AttrDict(enable=True, channel_wise_scale_type="momentum_act_max", momentum=0.95, **smooth)
    aq = AttrDict(n_bits=a_bits, channel_wise=False, per_group="token", scale_method="min_max",
                  round_mode="nearest_ste", running_stat=False, dynamic=True, sym=False, n_spatial_token=n_spatial,
                  n_temporal_token=n_temporal, n_prompt=n_prompt, smooth_quant=sq)
    return wq, aq


def install_opensora():
    """Make `opensora.models.stdit.stdit` / `opensora.models.layers.blocks` (the reference STDiT graph) importable
    without executing opensora's package __init__ files (they pull in diffusers / transformers / colossalai).
    Third-party pieces are replaced by minimal stand-ins: timm's DropPath / Mlp, and xformers' block-diagonal
    memory_efficient_attention restated with torch SDPA per (sample, prompt) segment."""
    install()
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    t2v = os.path.join(REFERENCE_ROOT, "t2v")

    def pkg(name, path):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [path]
            sys.modules[name] = m

    pkg("opensora", os.path.join(t2v, "opensora"))
    pkg("opensora.models", os.path.join(t2v, "opensora", "models"))
    pkg("opensora.models.layers", os.path.join(t2v, "opensora", "models", "layers"))
    pkg("opensora.models.stdit", os.path.join(t2v, "opensora", "models", "stdit"))
    pkg("opensora.acceleration", os.path.join(t2v, "opensora", "acceleration"))
    pkg("opensora.utils", os.path.join(t2v, "opensora", "utils"))
    if "opensora.registry" not in sys.modules:
        class _Registry:
            def register_module(self, *a, **k):
                def deco(obj):
                    return obj
                if len(a) == 1 and callable(a[0]) and not k:
                    return a[0]
                return deco
        _module("opensora.registry", MODELS=_Registry(), SCHEDULERS=_Registry())
    if "opensora.utils.ckpt_utils" not in sys.modules:
        _module("opensora.utils.ckpt_utils", load_checkpoint=lambda *a, **k: None)
    if "timm" not in sys.modules:
        class DropPath(nn.Identity):
            def __init__(self, *a, **k):
                super().__init__()

        class Mlp(nn.Module):
            def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0,
                         **unused):
                super().__init__()
                self.fc1 = nn.Linear(in_features, hidden_features or in_features)
                self.act = act_layer()
                self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)

            def forward(self, x):
                return self.fc2(self.act(self.fc1(x)))
        _module("timm")
        _module("timm.models")
        _module("timm.models.layers", DropPath=DropPath)
        _module("timm.models.vision_transformer", Mlp=Mlp)
    if "xformers" not in sys.modules:
        class BlockDiagonalMask:
            def __init__(self, q_lens, kv_lens):
                self.q_lens, self.kv_lens = list(q_lens), list(kv_lens)

            @classmethod
            def from_seqlens(cls, q_seqlen, kv_seqlen=None):
                return cls(q_seqlen, kv_seqlen if kv_seqlen is not None else q_seqlen)

        def memory_efficient_attention(q, k, v, p=0.0, attn_bias=None):
            # q [1, sum(Nq), H, D], k/v [1, sum(Nk), H, D]
            if attn_bias is None:
                o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
                return o.transpose(1, 2)
            outs, qo, ko = [], 0, 0
            for nq, nk in zip(attn_bias.q_lens, attn_bias.kv_lens):
                qq = q[:, qo:qo + nq].transpose(1, 2)
                kk = k[:, ko:ko + nk].transpose(1, 2)
                vv = v[:, ko:ko + nk].transpose(1, 2)
                outs.append(F.scaled_dot_product_attention(qq, kk, vv).transpose(1, 2))
                qo += nq
                ko += nk
            return torch.cat(outs, dim=1)
        fmha = _module("xformers.ops.fmha", BlockDiagonalMask=BlockDiagonalMask)
        xo = _module("xformers.ops", memory_efficient_attention=memory_efficient_attention, fmha=fmha)
        _module("xformers", ops=xo)


def install_pixart():
    """Make `diffusion.model.nets.PixArtMS` (reference t2i graph) importable: path-only packages (no __init__
    execution), a no-op mmcv-style registry, and constructor-only restatements of the timm pieces it subclasses."""
    install_opensora()   # omegaconf / diffusers / xformers / qdiff stubs (shared)
    import torch
    import torch.nn as nn
    t2i = os.path.join(REFERENCE_ROOT, "t2i")

    def pkg(name, path):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [path]
            sys.modules[name] = m

    pkg("diffusion", os.path.join(t2i, "diffusion"))
    pkg("diffusion.model", os.path.join(t2i, "diffusion", "model"))
    pkg("diffusion.model.nets", os.path.join(t2i, "diffusion", "model", "nets"))
    pkg("diffusion.utils", os.path.join(t2i, "diffusion", "utils"))
    if "diffusion.model.builder" not in sys.modules:
        class _Registry:
            def register_module(self, *a, **k):
                if len(a) == 1 and callable(a[0]) and not k:
                    return a[0]
                return lambda obj: obj
        _module("diffusion.model.builder", MODELS=_Registry())
    if "diffusion.model.utils" not in sys.modules:
        def auto_grad_checkpoint(module, *args, **kwargs):
            return module(*args, **kwargs)

        def to_2tuple(x):
            return tuple(x) if isinstance(x, (tuple, list)) else (x, x)
        _module("diffusion.model.utils", auto_grad_checkpoint=auto_grad_checkpoint, to_2tuple=to_2tuple,
                set_grad_checkpoint=lambda *a, **k: None)
    if "diffusion.utils.logger" not in sys.modules:
        import logging
        _module("diffusion.utils.logger", get_root_logger=lambda *a, **k: logging.getLogger("pixart"))
    vt = sys.modules["timm.models.vision_transformer"]
    if not hasattr(vt, "Attention"):
        class Attention(nn.Module):
            def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0.0, proj_drop=0.0, **unused):
                super().__init__()
                self.num_heads = num_heads
                self.scale = (dim // num_heads) ** -0.5
                self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
                self.attn_drop = nn.Dropout(attn_drop)
                self.proj = nn.Linear(dim, dim)
                self.proj_drop = nn.Dropout(proj_drop)

        class PatchEmbed(nn.Module):
            def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, bias=True, **unused):
                super().__init__()
                self.patch_size = (patch_size, patch_size)
                self.num_patches = (img_size // patch_size) ** 2
                self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)

            def forward(self, x):
                return self.proj(x).flatten(2).transpose(1, 2)
        vt.Attention = Attention
        vt.PatchEmbed = PatchEmbed
