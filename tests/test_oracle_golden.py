"""Pin oracle/qdiff_oracle.py against vectors produced by the unmodified reference (tests/golden/qdiff_golden.npz).

Bit-exact for everything the quantiser defines (delta, zero point, integer codes, fake-quantised tensors); the fp16
F.linear output is compared within fp16 accumulation-order noise (the reference's CPU half GEMM and our fp32-accumulate
restatement may round the last bit differently).
"""
import numpy as np
import pytest

from oracle import qdiff_oracle as O


def _cases(golden, prefix):
    return sorted(k for k in golden if k.startswith(prefix))


def test_golden_has_expected_cases(golden):
    assert len(_cases(golden, "act/")) >= 8
    assert len(_cases(golden, "layer/")) >= 14


@pytest.mark.parametrize("name", ["act/basic", "act/pooled_b2", "act/heavy_tail", "act/ragged_kv", "act/fc2_k4608",
                                  "act/signs_and_ranges", "act/bits6", "act/bits4"])
def test_act_quantizer_bit_exact(golden, name):
    c = golden[name]
    r = O.dynamic_act_quant(c["x"], int(c["n_bits"]))
    assert not r["degenerate"]
    np.testing.assert_array_equal(r["delta"], c["delta"].astype(np.float32))
    np.testing.assert_array_equal(r["zp"], c["zp"].astype(np.float32))
    np.testing.assert_array_equal(r["codes"], c["codes"])
    np.testing.assert_array_equal(r["xhat"], c["xhat"])
    np.testing.assert_array_equal(r["rowsum"], c["codes"].astype(np.int64).sum(-1))


def _layer_inputs(c):
    """Re-derive what each reference subclass feeds its act quantiser (the view tricks of stdit_quant_layer.py:70,161)."""
    x = c["x"]
    smooth = None
    if "act_scale" in c:
        tr = c["timerange"]
        t = int(c["t_eval"])
        idx = next(i for i, (lo, hi) in enumerate(tr) if lo <= t <= hi)
        smooth = O.smooth_channel_scale(c["act_scale"][idx].reshape(-1), c["weight"], float(c["alpha"][idx]))
    return x, smooth


LAYER_VIEWS = {
    # name -> (pool batch B, tokens) given x.shape; spatial/temporal layers view (B*T,S,C)->(B,T*S,C) with T=4,S=16
    "layer/mlp_fc1": lambda s: (s[0], s[1]),
    "layer/mlp_fc2_k4608": lambda s: (s[0], s[1]),
    "layer/spatial_attn": lambda s: (s[0] // 4, 4 * s[1]),
    "layer/spatial_attn_b2": lambda s: (s[0] // 4, 4 * s[1]),
    "layer/temporal_attn": lambda s: (s[0] // 16, 16 * s[1]),
    "layer/cross_q": lambda s: (s[0], s[1]),
    "layer/cross_kv": lambda s: (s[0], s[1]),
    "layer/pixart_qkv_b2": lambda s: (s[0], s[1]),
    "layer/pixart_cross_kv": lambda s: (s[0], s[1]),
    "layer/nobias": lambda s: (s[0], s[1]),
    "layer/w4_plain": lambda s: (s[0], s[1]),
    "layer/w4_smooth_t100": lambda s: (s[0] // 4, 4 * s[1]),
    "layer/w4_smooth_t900": lambda s: (s[0] // 4, 4 * s[1]),
    "layer/w8_smooth_mlp_t700": lambda s: (s[0], s[1]),
}


@pytest.mark.parametrize("name", sorted(LAYER_VIEWS))
def test_quant_layer_family(golden, name):
    c = golden[name]
    x, smooth = _layer_inputs(c)
    B, n = LAYER_VIEWS[name](x.shape)
    xv = x.reshape(B, n, x.shape[-1])
    a = O.dynamic_act_quant(xv, 8, smooth)
    np.testing.assert_array_equal(a["delta"], c["adelta"].astype(np.float32))
    np.testing.assert_array_equal(a["zp"], c["azp"].astype(np.float32))
    w_bits = int(c["w_bits"])
    bias = c["bias"] if "bias" in c else None
    fake = O.quant_linear_fake(xv, c["weight"], bias, c["wdelta"], c["wzp"], w_bits, 8, smooth)
    ref = c["out"].reshape(B, n, -1).astype(np.float32)
    diff = np.abs(fake.astype(np.float32) - ref)
    # same fake-quant operands, different fp16 GEMM accumulation order: at most a couple of fp16 ulps
    assert diff.max() <= 4 * np.spacing(np.abs(ref).max().astype(np.float16)).astype(np.float32)
    assert (diff > 0).mean() < 0.35
    # integer decomposition (what the CUDA kernel evaluates) against the reference's fake-quant output: <= 1e-3 relative
    wq = O.weight_quant(c["weight"], c["wdelta"], c["wzp"], w_bits, smooth)
    assert wq["codes"].max() <= 2 ** w_bits - 1
    y = O.quant_linear_int(a["codes"], a["delta"], a["zp"], a["rowsum"], wq["codes"], c["wdelta"], c["wzp"], bias)
    rel_inf = np.abs(y.astype(np.float32) - ref).max() / np.abs(ref).max()
    rel_l2 = np.linalg.norm(y.astype(np.float32) - ref) / np.linalg.norm(ref)
    assert rel_inf <= 1e-3 and rel_l2 <= 1e-3, (rel_inf, rel_l2)


def test_smooth_quant_uses_timerange0_weight_grid(golden):
    """Quirk Q7: after init_done the weight delta/zero-point never switch timerange (base_quantizer.py:114-127)."""
    a, b = golden["layer/w4_smooth_t100"], golden["layer/w4_smooth_t900"]
    np.testing.assert_array_equal(a["wdelta"], b["wdelta"])
    np.testing.assert_array_equal(a["wzp"], b["wzp"])
    assert not np.array_equal(a["out"], b["out"])  # but the channel scale (alpha, act_scale) does switch


def test_eps_quirk_flags_degenerate():
    """Quirk Q4 (base_quantizer.py:220-223): one ~zero-range token makes delta.fill_(1e-6) hit every row."""
    x = np.random.default_rng(0).standard_normal((1, 8, 64)).astype(np.float16)
    x[0, 3] = 0
    r = O.dynamic_act_quant(x, 8)
    assert r["degenerate"]
    assert np.all(r["delta"] == np.float32(np.float16(1e-6)))


def test_div_by_levels_matches_reciprocal_multiply():
    """The CUDA build of torch divides a half tensor by a Python scalar as a * (1/b) in fp32; the CPU build divides.
    Both round to the same fp16 for every positive half and every level count the configs use, so the pinned CPU
    behaviour is also the GPU reference's."""
    allh = np.arange(0, 0x7C00, dtype=np.uint16).view(np.float16).astype(np.float32)
    for qmax in (15.0, 63.0, 255.0):
        a = (allh / np.float32(qmax)).astype(np.float16)
        b = (allh * (np.float32(1.0) / np.float32(qmax))).astype(np.float16)
        mism = int((a != b).sum())
        assert mism == 0, (qmax, mism)


@pytest.mark.parametrize("name", sorted(LAYER_VIEWS))
def test_torch_fake_quant_restatement_is_bit_exact_on_cpu(golden, name):
    """oracle/torch_fake_quant.py (used on the GPU by the model-level tests) reproduces the reference layer outputs
    exactly when run on the reference's own back end (CPU, fp16)."""
    import torch
    from oracle import torch_fake_quant as TF
    c = golden[name]
    x, smooth = _layer_inputs(c)
    B, n = LAYER_VIEWS[name](x.shape)
    xv = torch.from_numpy(x.reshape(B, n, x.shape[-1]))
    sm = None if smooth is None else torch.from_numpy(smooth)
    bias = torch.from_numpy(c["bias"]) if "bias" in c else None
    out = TF.quant_linear_fake(xv, torch.from_numpy(c["weight"]), bias, torch.from_numpy(c["wdelta"]),
                               torch.from_numpy(c["wzp"]), int(c["w_bits"]), 8, sm)
    np.testing.assert_array_equal(out.numpy().reshape(c["out"].shape), c["out"])


@pytest.mark.parametrize("name", ["layer/mlp_fc1", "layer/spatial_attn_b2", "layer/w4_smooth_t900", "layer/cross_kv"])
def test_torch_exact_operand_simulation_equals_integer_decomposition(golden, name):
    """oracle.torch_fake_quant(exact=True) — the simulation with un-rounded dequantised operands — is the same number
    as the integer decomposition the kernels evaluate (up to fp32 summation order)."""
    import torch
    from oracle import torch_fake_quant as TF
    c = golden[name]
    x, smooth = _layer_inputs(c)
    B, n = LAYER_VIEWS[name](x.shape)
    xv = x.reshape(B, n, x.shape[-1])
    sm = None if smooth is None else torch.from_numpy(smooth)
    bias = c["bias"] if "bias" in c else None
    ex = TF.quant_linear_fake(torch.from_numpy(xv), torch.from_numpy(c["weight"]),
                              None if bias is None else torch.from_numpy(bias), torch.from_numpy(c["wdelta"]),
                              torch.from_numpy(c["wzp"]), int(c["w_bits"]), 8, sm, exact=True).numpy()
    a = O.dynamic_act_quant(xv, 8, smooth)
    wq = O.weight_quant(c["weight"], c["wdelta"], c["wzp"], int(c["w_bits"]), smooth)
    y = O.quant_linear_int(a["codes"], a["delta"], a["zp"], a["rowsum"], wq["codes"], c["wdelta"], c["wzp"], bias)
    d = np.abs(ex.astype(np.float32) - y.astype(np.float32))
    assert d.max() <= 2 * float(np.spacing(np.float16(np.abs(y.astype(np.float32)).max())))
    assert (d > 0).mean() < 0.02


STATIC_CASES = ["static/tensor_mlp", "static/tensor_spatial_b2", "static/tensor_bits6", "static/token_mlp"]


def _static_view(c):
    """[B, n, C] view whose dim 1 is the token axis of a static per-token quantiser (plain QuantLayer: as given)."""
    x = c["x"]
    return x.reshape(1, -1, x.shape[-1]) if not int(c["per_token"]) else x


@pytest.mark.parametrize("name", STATIC_CASES)
def test_static_act_quantizer_bit_exact(golden_static, name):
    """N3 / w8a8_naive.yaml: calibrated per-tensor (and static per-token) activation scales, codes pinned bit-for-bit
    to the unmodified reference, output of the whole layer within 1e-3 through the integer decomposition."""
    c = golden_static[name]
    bits = int(c["a_bits"])
    xv = _static_view(c)
    r = O.static_act_quant(xv, c["adelta"], c["azp"], bits)
    np.testing.assert_array_equal(r["codes"].reshape(c["codes"].shape), c["codes"])
    assert (c["codes"] == 0).any() and (c["codes"] == 2 ** bits - 1).any()      # the vectors do saturate
    B, n, K = xv.shape
    d = np.broadcast_to(c["adelta"].astype(np.float32), (n,)) if c["adelta"].size == 1 else c["adelta"].astype(np.float32)
    z = np.broadcast_to(c["azp"].astype(np.float32), (n,)) if c["azp"].size == 1 else c["azp"].astype(np.float32)
    wq = O.weight_quant(c["weight"], c["wdelta"], c["wzp"], 8)
    y = O.quant_linear_int(r["codes"], d, z, r["rowsum"], wq["codes"], c["wdelta"], c["wzp"], c["bias"])
    ref = c["out"].reshape(B, n, -1).astype(np.float32)
    rel_inf = np.abs(y.astype(np.float32) - ref).max() / np.abs(ref).max()
    rel_l2 = np.linalg.norm(y.astype(np.float32) - ref) / np.linalg.norm(ref)
    assert rel_inf <= 1e-3 and rel_l2 <= 1e-3, (rel_inf, rel_l2)
