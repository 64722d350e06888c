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
