"""One-call switch for a user of the reference: `from_reference(ref_qnn)` takes an UNMODIFIED reference `QuantModel`
(qdiff/models/quant_model.py:38) wrapped around the reference `STDiT` (t2v/opensora/models/stdit/stdit.py) or `PixArtMS`
(t2i/diffusion/model/nets/PixArtMS.py) — weights loaded, PTQ checkpoint loaded or calibrated, layer states set the way the
inference scripts set them (quant_txt2video.py:195-207) — and returns this repo's `(QuantModel, model)` pair carrying the
same weights, quantiser parameters and layer states, whose `model.forward_fused` is the block-fused B200 schedule
(`bench.py`'s 40 ms step; the reference objects themselves can only take the layer-by-layer `accelerate()` path, 3x slower).

Nothing of the reference is imported here: the reference objects are read by attribute name only (the module tree of
`viditq_b200.stdit.STDiT` / `viditq_b200.pixart.PixArtMS` mirrors the reference's names, so state_dict keys coincide).
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from .qdiff import _LAYER_STATE, QuantModel


def _is_ref_quant_layer(m):
    return hasattr(m, "org_module") and hasattr(m, "weight_quantizer") and hasattr(m, "act_quantizer")


def plain_state_dict(ref_model):
    """The reference model's weights under their ORIGINAL names: a reference QuantLayer registers its wrapped layer as
    `<name>.org_module.*` (and aliases `.weight` / `.org_weight`); everything else keeps its key."""
    out = OrderedDict()
    wrapped = []
    for name, m in ref_model.named_modules():
        if _is_ref_quant_layer(m):
            wrapped.append(name + ".")
            for pname, p in m.org_module.named_parameters():
                out[f"{name}.{pname}"] = p.detach()
    for key, v in ref_model.state_dict().items():
        if not any(key.startswith(w) for w in wrapped):
            out[key] = v.detach()
    return out


def _build_like(ref):
    kind = type(ref).__name__
    y_proj = ref.y_embedder.y_proj
    fc1 = y_proj.fc1.org_module if _is_ref_quant_layer(y_proj.fc1) else y_proj.fc1
    # attributes the reference classes do not all keep (PixArtMS stores neither hidden_size nor mlp_ratio) come from shapes
    hidden = getattr(ref, "hidden_size", None) or ref.blocks[0].scale_shift_table.shape[-1]
    fc1_mlp = ref.blocks[0].mlp.fc1
    fc1_mlp = fc1_mlp.org_module if _is_ref_quant_layer(fc1_mlp) else fc1_mlp
    common = dict(in_channels=ref.in_channels, hidden_size=hidden, depth=getattr(ref, "depth", len(ref.blocks)),
                  num_heads=getattr(ref, "num_heads", None) or ref.blocks[0].attn.num_heads,
                  mlp_ratio=getattr(ref, "mlp_ratio", None) or fc1_mlp.out_features / hidden,
                  pred_sigma=ref.out_channels == 2 * ref.in_channels, caption_channels=fc1.in_features,
                  model_max_length=ref.y_embedder.y_embedding.shape[0], dtype=getattr(ref, "dtype", torch.float32))
    if kind.startswith("STDiT"):
        from .stdit import STDiT
        if getattr(ref, "no_temporal_pos_emb", False):
            raise NotImplementedError("from_reference: STDiT with no_temporal_pos_emb")
        return STDiT(input_size=tuple(ref.input_size), patch_size=tuple(ref.patch_size),
                     space_scale=getattr(ref, "space_scale", 1.0), time_scale=getattr(ref, "time_scale", 1.0), **common), "opensora"
    if kind.startswith("PixArt"):
        from .pixart import PixArtMS
        ps = ref.patch_size if isinstance(ref.patch_size, int) else ref.patch_size[0]
        return PixArtMS(input_size=ref.base_size * ps, patch_size=ps,
                        pe_interpolation=getattr(ref, "pe_interpolation", getattr(ref, "lewei_scale", 1.0)), **common), "pixart"
    raise NotImplementedError(f"from_reference: {kind} (STDiT and PixArtMS are the quantised model families)")


@torch.no_grad()
def from_reference(ref_qnn):
    """-> (qnn, model): viditq_b200.qdiff.QuantModel around viditq_b200.stdit.STDiT / pixart.PixArtMS, on the device and
    in the dtype of the reference model, carrying its weights, every quantiser buffer (the ckpt.pth content) and every
    layer-level switch.  The reference objects are left untouched."""
    ref = ref_qnn.model
    model, model_type = _build_like(ref)
    sd = plain_state_dict(ref)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    # buffers this repo recomputes (sin-cos tables) may be absent from / extra in the reference: parameters must all match
    params = {n for n, _ in model.named_parameters()}
    bad = [k for k in missing if k in params] + [k for k in unexpected if k.rsplit(".", 1)[-1] in ("weight", "bias")]
    if bad:
        raise RuntimeError(f"from_reference: parameters do not line up with the reference model: {bad[:6]}")
    model.eval()
    any_layer = next(m for m in ref.modules() if _is_ref_quant_layer(m))
    qnn = QuantModel(model, any_layer.weight_quant_params, any_layer.act_quant_params, model_type=model_type)
    p = next(ref.parameters())
    qnn.to(p.device)
    if p.dtype == torch.float16:
        qnn.half()
        model.dtype = torch.float16
    ref_qnn.set_module_name_for_quantizer(module=ref_qnn.model)
    qnn.set_module_name_for_quantizer(module=qnn.model)
    ckpt = ref_qnn.get_quant_params_dict()
    qnn.set_quant_params_dict(ckpt, dtype=p.dtype)
    mine = dict(qnn.quant_layers())
    for name, m in ref.named_modules():
        if not _is_ref_quant_layer(m):
            continue
        ours = mine.get(name)
        if ours is None:
            if isinstance(m.org_module, nn.Linear):
                raise RuntimeError(f"from_reference: no counterpart for the reference QuantLayer {name}")
            continue                       # Conv QuantLayers (PixArt's x_embedder): plain convolutions here, FP in every script
        for attr in _LAYER_STATE:
            if hasattr(m, attr):
                setattr(ours, attr, getattr(m, attr))
        for qname in ("weight_quantizer", "act_quantizer"):
            rq, oq = getattr(m, qname), getattr(ours, qname)
            for attr in ("n_bits", "bit_idx", "init_done", "cur_timestep_id", "x_min", "x_max"):
                if hasattr(rq, attr):
                    setattr(oq, attr, getattr(rq, attr))
        ours.invalidate_prepared()
    for attr in ("cfg_split", "timestep_wise", "fp_layer_list", "smooth_quant_stat", "timestep_wise_mp",
                 "time_mp_config_weight", "time_mp_config_act"):
        if attr in ref_qnn.__dict__:
            setattr(qnn, attr, ref_qnn.__dict__[attr])
    return qnn, model
