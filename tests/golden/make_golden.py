"""Generate tests/golden/qdiff_golden.npz by executing the UNMODIFIED reference classes (thu-nics/ViDiT-Q @ 44126bd,
/root/reference) on seeded inputs, in fp16 on CPU (the dtype the reference runs its model and quantiser buffers in).

Run here (the container with /root/reference):   python tests/golden/make_golden.py
The reference cannot travel to the GPU box; the committed .npz is what the GPU parity tests and the oracle pin against.
Flow per layer mirrors t2v/scripts/ptq.py:266-294 (fp32 weight-init forward with weight_quant on, act_quant off, then
set_quant_init_done) followed by quant_txt2video.py:195-207 (quant state on, `.to(fp16)`).
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
from qdiff.quantizer.dynamic_quantizer import DynamicActQuantizer  # noqa: E402
from qdiff.models.quant_layer import QuantLayer  # noqa: E402
from qdiff.models.stdit_quant_layer import (QuantSpatialAttnLinear, QuantTemporalAttnLinear,  # noqa: E402
                                            QuantCrossAttnLinear)
from qdiff.models.dit_quant_layer import QuantAttnLinearImg, QuantCrossAttnLinearImg  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "qdiff_golden.npz")
G = {}


def put(name, **arrs):
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        G[f"{name}/{k}"] = np.asarray(v)


def act_case(name, x, n_bits=8):
    _, aq = ref_shims.w8a8_dynamic_configs(a_bits=n_bits)
    q = DynamicActQuantizer(aq)
    q.init_done = True
    q.module_name = name
    xhat = q(x)
    codes = q.rounding(x)  # the reference's own (dead, dynamic_quantizer.py:43) integer-code routine
    put(name, x=x, n_bits=n_bits, delta=q.delta.reshape(-1), zp=q.zero_point.reshape(-1), codes=codes.to(torch.uint8),
        xhat=xhat)


def make_layer(cls, C, N, T, S, n_prompt, w_bits, seed, smooth=None, bias=True):
    """Build a reference layer the way QuantModel + ptq.py + quant_txt2video.py do, return it in fp16 inference state."""
    torch.manual_seed(seed)
    lin = nn.Linear(C, N, bias=bias)
    with torch.no_grad():
        lin.weight.mul_(1.5)
        if bias:
            lin.bias.normal_(0, 0.02)
    wq, aq = ref_shims.w8a8_dynamic_configs(n_temporal=T, n_spatial=S, n_prompt=n_prompt, w_bits=w_bits, smooth=smooth)
    wq["mixed_precision"] = [4, 6, 8]
    layer = cls(lin, wq, aq)
    layer.weight_quantizer.module_name = "w"
    layer.act_quantizer.module_name = "a"
    layer.cur_timestep_id = 0
    act_scale = None
    if smooth is not None:
        layer.smooth_quant = True
        act_scale = torch.randn(len(smooth["timerange"]), 1, C).abs() + 0.5
        layer.act_quantizer.act_scale = act_scale.clone()
    layer.set_quant_state(True, False)
    return layer, act_scale


def layer_case(name, cls, xshape, C=1152, N=128, T=4, S=16, n_prompt=8, w_bits=8, seed=0, smooth=None, t_eval=0,
               bias=True, heavy=False):
    layer, act_scale = make_layer(cls, C, N, T, S, n_prompt, w_bits, seed, smooth, bias)
    torch.manual_seed(seed + 1000)
    x32 = torch.randn(*xshape)
    if heavy:
        x32[..., ::97] *= 25.0
    # weight init pass(es) in fp32 (ptq.py:276-291): one forward per timerange start when smooth-quant is timerange-aware
    starts = [r[0] for r in smooth["timerange"]] if smooth is not None else [0]
    for st in starts:
        layer.cur_timestep_id = st
        _ = layer(x32)
    layer.weight_quantizer.init_done = True
    layer.act_quantizer.init_done = True
    layer.set_quant_state(True, True)
    layer.half()
    layer.cur_timestep_id = t_eval
    x16 = x32.half()
    with torch.no_grad():
        out = layer(x16)
    wqz = layer.weight_quantizer
    rec = dict(x=x16, weight=layer.weight, out=out, wdelta=wqz.delta.reshape(-1), wzp=wqz.zero_point.reshape(-1),
               w_bits=wqz.n_bits, t_eval=t_eval, adelta=layer.act_quantizer.delta.reshape(-1),
               azp=layer.act_quantizer.zero_point.reshape(-1))
    if bias:
        rec["bias"] = layer.bias
    if smooth is not None:
        rec["act_scale"] = layer.act_quantizer.act_scale  # fp16 after .half()
        rec["alpha"] = np.asarray(smooth["alpha"], dtype=np.float64)
        rec["timerange"] = np.asarray(smooth["timerange"], dtype=np.int64)
    put(name, **rec)


def main():
    torch.set_grad_enabled(False)
    torch.manual_seed(0)
    # ---- a1: DynamicActQuantizer -------------------------------------------------------------------------------
    act_case("act/basic", torch.randn(1, 64, 1152).half())
    x = torch.randn(2, 48, 1152)
    x[1] *= 3.0
    act_case("act/pooled_b2", x.half())                       # quirk Q1: statistics pooled over the batch
    x = torch.randn(1, 40, 1152)
    x[..., 7] *= 30.0
    x[..., 500] *= -18.0
    act_case("act/heavy_tail", x.half())
    act_case("act/ragged_kv", (torch.randn(1, 37, 1152) * 0.3).half())   # cross-attn kv: mask-selected prompt tokens
    act_case("act/fc2_k4608", torch.nn.functional.gelu(torch.randn(1, 12, 4608), approximate="tanh").half())
    x = torch.randn(1, 16, 1152)
    x[0, 0] = x[0, 0].abs() + 0.1      # all-positive row -> min clamps to 0
    x[0, 1] = -x[0, 1].abs() - 0.1     # all-negative row -> max clamps to 0
    x[0, 2] *= 1e-3                    # small (not degenerate) range
    x[0, 3] *= 300.0                   # large magnitudes
    act_case("act/signs_and_ranges", x.half())
    act_case("act/bits6", torch.randn(1, 24, 1152).half(), n_bits=6)
    act_case("act/bits4", torch.randn(1, 24, 1152).half(), n_bits=4)

    # ---- a3-a7: QuantLayer family ------------------------------------------------------------------------------
    T, S = 4, 16
    layer_case("layer/mlp_fc1", QuantLayer, (1, T * S, 1152), N=256, seed=1)
    layer_case("layer/mlp_fc2_k4608", QuantLayer, (1, 24, 4608), C=4608, N=64, seed=2)
    layer_case("layer/spatial_attn", QuantSpatialAttnLinear, (1 * T, S, 1152), seed=3)
    layer_case("layer/spatial_attn_b2", QuantSpatialAttnLinear, (2 * T, S, 1152), seed=4)
    layer_case("layer/temporal_attn", QuantTemporalAttnLinear, (1 * S, T, 1152), seed=5)
    layer_case("layer/cross_q", QuantCrossAttnLinear, (1, T * S, 1152), seed=6, heavy=True)
    layer_case("layer/cross_kv", QuantCrossAttnLinear, (1, 13, 1152), N=256, seed=7)
    layer_case("layer/pixart_qkv_b2", QuantAttnLinearImg, (2, 40, 1152), N=384, seed=8)
    layer_case("layer/pixart_cross_kv", QuantCrossAttnLinearImg, (1, 21, 1152), N=256, seed=9)
    layer_case("layer/nobias", QuantLayer, (1, 32, 1152), N=64, seed=10, bias=False)
    # W4 weights (w4a8): same kernel, 4-bit grid
    layer_case("layer/w4_plain", QuantLayer, (1, 32, 1152), N=128, w_bits=4, seed=11)
    # timerange-aware smooth quant (w4a8_timestep_aware_cb.yaml), evaluated in each timerange (quirk Q7: timerange-0 delta)
    sq = dict(alpha=[0.11, 0.31], timerange=[[0, 500], [501, 1000]])
    layer_case("layer/w4_smooth_t100", QuantSpatialAttnLinear, (1 * T, S, 1152), w_bits=4, seed=12, smooth=sq,
               t_eval=100)
    layer_case("layer/w4_smooth_t900", QuantSpatialAttnLinear, (1 * T, S, 1152), w_bits=4, seed=12, smooth=sq,
               t_eval=900)
    layer_case("layer/w8_smooth_mlp_t700", QuantLayer, (1, 32, 1152), N=64, w_bits=8, seed=13, smooth=sq, t_eval=700)

    np.savez_compressed(OUT, **G)
    print(f"wrote {OUT}: {len(G)} arrays, {os.path.getsize(OUT) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
