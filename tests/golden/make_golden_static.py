"""Generate tests/golden/static_act_golden.npz: the reference's STATIC activation-quantiser path (w8a8_naive.yaml:
`per_group: False, dynamic: False` — one calibrated (delta, zero_point) per tensor; plus the static per-token variant),
by executing the UNMODIFIED reference classes from /root/reference in fp16 on CPU.

    python tests/golden/make_golden_static.py

Flow mirrors ptq.py: fp32 calibration forwards with weight_quant / act_quant on and init_done False (the quantisers
initialise themselves from the calibration tensors, base_quantizer.py:114-127), then set_quant_init_done, `.half()`, and
an inference forward on a DIFFERENT tensor (whose range exceeds the calibrated one, so codes saturate).
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
from qdiff.models.quant_layer import QuantLayer  # noqa: E402
from qdiff.models.stdit_quant_layer import QuantSpatialAttnLinear  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "static_act_golden.npz")
G = {}


def put(name, **arrs):
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        G[f"{name}/{k}"] = np.asarray(v)


def case(name, cls, xshape, per_group, C=1152, N=96, T=4, S=16, a_bits=8, seed=0):
    torch.manual_seed(seed)
    lin = nn.Linear(C, N)
    with torch.no_grad():
        lin.weight.mul_(1.5)
        lin.bias.normal_(0, 0.02)
    wq, aq = ref_shims.w8a8_dynamic_configs(n_temporal=T, n_spatial=S, n_prompt=8, a_bits=a_bits)
    aq["dynamic"] = False
    aq["per_group"] = per_group          # False = tensor-wise (w8a8_naive.yaml) | "token" = static per-token
    layer = cls(lin, wq, aq)
    layer.weight_quantizer.module_name = "w"
    layer.act_quantizer.module_name = "a"
    layer.cur_timestep_id = 0
    layer.set_quant_state(True, True)
    x_cal = torch.randn(*xshape)
    _ = layer(x_cal)                      # calibration: both quantisers initialise from this call
    layer.weight_quantizer.init_done = True
    layer.act_quantizer.init_done = True
    layer.half()
    x = (torch.randn(*xshape) * 1.6).half()   # wider than the calibration tensor: saturating codes
    x[..., 11] *= 9.0
    with torch.no_grad():
        out = layer(x)
    aqz, wqz = layer.act_quantizer, layer.weight_quantizer
    # the integer codes behind the reference's x_dequant (base_quantizer.py:134-143), recomputed with the same fp16 ops
    # and checked against the module's own output of the quantiser
    codes = torch.clamp(torch.round(x / aqz.delta) + aqz.zero_point, 0, 2 ** a_bits - 1)
    assert torch.equal((codes - aqz.zero_point) * aqz.delta, aqz(x))
    put(name, x=x, weight=layer.weight, bias=layer.bias, out=out, wdelta=wqz.delta.reshape(-1),
        wzp=wqz.zero_point.reshape(-1), adelta=aqz.delta.reshape(-1), azp=aqz.zero_point.reshape(-1),
        codes=codes.to(torch.uint8), a_bits=a_bits, per_token=int(per_group == "token"), T=T, S=S)


def main():
    torch.set_grad_enabled(False)
    case("static/tensor_mlp", QuantLayer, (1, 64, 1152), False, seed=21)
    case("static/tensor_spatial_b2", QuantSpatialAttnLinear, (2 * 4, 16, 1152), False, seed=22)
    case("static/tensor_bits6", QuantLayer, (1, 40, 1152), False, a_bits=6, seed=23)
    case("static/token_mlp", QuantLayer, (2, 48, 1152), "token", seed=24)
    np.savez_compressed(OUT, **G)
    print(f"wrote {OUT}: {len(G)} arrays, {os.path.getsize(OUT) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
