"""PixArt-alpha/MS (BASELINE config 2 family): host graph pinned on CPU against the reference's
fp32 output, and — on the GPU — the W8A8 integer kernels against the simulated quantisation (same back end) and the
reference's own fp16 W8A8 output (tests/golden/pixart_small_golden.npz, generator make_golden_pixart.py)."""
import os

import numpy as np
import pytest
import torch

from test_stdit_graph_cpu import Cfg, ckpt_from_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP_LAYERS = ["x_embedder", "t_embedder", "t_block", "y_embedder", "csize_embedder", "ar_embedder"]


@pytest.fixture(scope="module")
def pix():
    z = np.load(os.path.join(ROOT, "tests", "golden", "pixart_small_golden.npz"))
    return {k: z[k] for k in z.files}


def build(pix):
    from viditq_b200.pixart import PixArtMS
    from viditq_b200.qdiff import QuantModel
    model = PixArtMS(input_size=16, depth=2)
    model.init_synthetic(seed=0)
    model.eval()
    sq = Cfg(enable=False, channel_wise_scale_type="momentum_act_max", momentum=0.95, alpha=0.625)
    wq = Cfg(n_bits=8, per_group="channel", channel_dim=0, scale_method="min_max", round_mode="nearest",
             mixed_precision=[4, 6, 8])
    aq = Cfg(n_bits=8, per_group="token", scale_method="min_max", round_mode="nearest_ste", running_stat=False,
             dynamic=True, sym=False, n_spatial_token=64, n_temporal_token=1, n_prompt=120, smooth_quant=sq)
    qnn = QuantModel(model, wq, aq, model_type="pixart")
    qnn.set_module_name_for_quantizer(module=qnn.model)
    qnn.set_quant_params_dict(ckpt_from_golden(pix))
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    return qnn, model


def _rel(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return np.abs(a - b).max() / np.abs(b).max(), np.linalg.norm(a - b) / np.linalg.norm(b)


def test_pixart_layer_types_and_fp_graph(pix):
    from viditq_b200 import qdiff
    qnn, model = build(pix)
    b = model.blocks[0]
    assert type(b.attn.qkv) is qdiff.QuantAttnLinearImg and type(b.attn.proj) is qdiff.QuantAttnLinearImg
    assert type(b.cross_attn.kv_linear) is qdiff.QuantCrossAttnLinearImg
    assert type(b.mlp.fc2) is qdiff.QuantLayer and type(model.final_layer.linear) is qdiff.QuantLayer
    assert sum(1 for n, _ in qnn.quant_layers() if n.startswith("blocks.")) == 7 * 2   # 7 quantised linears per block
    qnn.set_quant_state(False, False)
    with torch.no_grad():
        out = qnn(torch.from_numpy(pix["x"]), torch.from_numpy(pix["t"]), torch.from_numpy(pix["y"]).float(),
                  mask=torch.from_numpy(pix["mask"])).numpy()
    inf, _ = _rel(out, pix["out_fp32"])
    assert inf < 2e-5, inf


@pytest.mark.gpu
def test_pixart_w8a8_on_gpu(pix):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import torch_fake_quant as TF
    from viditq_b200 import ops
    qnn, model = build(pix)
    qnn.cuda()
    qnn.half()
    model.dtype = torch.float16
    qnn.set_quant_state(True, True)
    qnn.set_layer_quant(model=qnn, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")
    assert model.final_layer.linear.get_quant_state() == (True, True)      # quantised in PixArt (unlike STDiT)
    x, t = torch.from_numpy(pix["x"]).cuda(), torch.from_numpy(pix["t"]).cuda()
    y, mask = torch.from_numpy(pix["y"]).cuda(), torch.from_numpy(pix["mask"]).cuda()

    saved = {}

    def sim_forward(exact):
        def make(layer):
            def fwd(inp, *a, **k):
                if not (layer.weight_quant and layer.act_quant):
                    return saved[layer](inp)
                wq = layer.weight_quantizer
                return TF.quant_linear_fake(inp, layer.weight, layer.bias, wq.delta, wq.zero_point, wq.n_bits, 8,
                                            exact=exact)
            return fwd
        for _, layer in qnn.quant_layers():
            saved[layer] = layer.forward
            layer.forward = make(layer)
        try:
            with torch.no_grad():
                return qnn(x, t, y, mask=mask).float().cpu().numpy()
        finally:
            for layer in saved:
                del layer.forward
    sim = sim_forward(True)            # same codes, un-rounded dequantised operands: what an integer kernel computes
    sim16 = sim_forward(False)         # the reference's fp16 simulation, on this back end
    band = _rel(sim16, sim)            # the reference's own noise band on this model
    # teacher-forced per-layer parity (the 1e-3 bar): each quantised linear on the activations it really sees
    errs = {}

    def probe(name, layer):
        def fwd(inp, *a, **k):
            out_ = saved[layer](inp)
            if layer.weight_quant and layer.act_quant:
                wq = layer.weight_quantizer
                ref = TF.quant_linear_fake(inp, layer.weight, layer.bias, wq.delta, wq.zero_point, wq.n_bits, 8)
                dd, rr = out_.float() - ref.float(), ref.float()
                errs[name] = ((dd.abs().max() / rr.abs().max()).item(), (dd.norm() / rr.norm()).item())
            return out_
        return fwd
    for name, layer in qnn.quant_layers():
        saved[layer] = layer.forward
        layer.forward = probe(name, layer)
    with torch.no_grad():
        qnn(x, t, y, mask=mask)
    for layer in saved:
        del layer.forward
    worst = max(errs.values(), key=lambda e: e[1])
    print("pixart per-layer parity inside the model (%d layers): worst rel-inf %.3e rel-L2 %.3e"
          % (len(errs), max(e[0] for e in errs.values()), worst[1]))
    assert len(errs) == 15 and worst[1] <= 1e-3 and max(e[0] for e in errs.values()) <= 1e-3
    n0 = ops.launch_count()
    with torch.no_grad():
        out = qnn(x, t, y, mask=mask).float().cpu().numpy()
        fused = model.forward_fused(x, t, y, mask=mask).float().cpu().numpy()
    assert ops.launch_count() - n0 >= 2 * 15 and ops.check_status() == 0     # >= one own launch per quantised linear, both schedules
    a, b = _rel(out, sim), _rel(fused, sim)
    c, d = _rel(out, pix["out_w8a8"]), _rel(fused, pix["out_w8a8"])
    print("pixart int vs exact-operand sim: layerwise %.3e %.3e | fused %.3e %.3e" % (a + b))
    print("pixart vs reference (CPU fp16) W8A8: layerwise %.3e %.3e | fused %.3e %.3e" % (c + d))
    qerr = _rel(pix["out_w8a8"], pix["out_fp16"])
    print("pixart reference noise band (fp16 sim vs exact-operand sim, this GPU): %.3e %.3e; quantisation error %.3e"
          % (band + (qerr[1],)))
    # End to end, the STDiT-style band test (tests/test_gpu_stdit.py): on this toy model the reference's OWN fp16
    # simulation sits 1.0e-2 from the same simulation with un-rounded operands (measured above; as large as the 8-bit
    # quantisation error, so a looser constant would prove nothing) — the integer kernels must be inside that band
    # around the simulation, and the fused schedule inside it around the layerwise one.  The strict statement is the
    # per-layer one asserted above (<= 1e-3 on all 15 quantised linears).
    assert a[1] <= band[1] and b[1] <= band[1], (a, b, band)
    assert c[1] <= 1.25 * band[1] and d[1] <= 1.25 * band[1], (c, d, band)
    assert _rel(fused, out)[1] <= band[1]


@pytest.mark.gpu
def test_pixart_512_fused_schedule_uses_the_attention_kernels():
    """BASELINE config 2 shape (PixArt 512x512: 64x64 latent -> 1024 tokens, CFG batch 2), two blocks of synthetic weights:
    the fused schedule (fused patch embedding, tcgen05 self- and cross-attention) against the layer-by-layer schedule
    (QuantLayer calls, Conv2d, torch SDPA) on the same GPU."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from viditq_b200 import ops
    from viditq_b200.pixart import PixArtMS
    from viditq_b200.qdiff import QuantModel
    model = PixArtMS(input_size=64, depth=2)
    model.init_synthetic(seed=2)
    model.eval()
    sq = Cfg(enable=False, channel_wise_scale_type="momentum_act_max", momentum=0.95, alpha=0.625)
    wq = Cfg(n_bits=8, per_group="channel", channel_dim=0, scale_method="min_max", round_mode="nearest",
             mixed_precision=[4, 6, 8])
    aq = Cfg(n_bits=8, per_group="token", scale_method="min_max", round_mode="nearest_ste", running_stat=False,
             dynamic=True, sym=False, n_spatial_token=1024, n_temporal_token=1, n_prompt=120, smooth_quant=sq)
    qnn = QuantModel(model, wq, aq, model_type="pixart")
    qnn.cuda()
    qnn.half()
    model.dtype = torch.float16
    qnn.set_module_name_for_quantizer(module=qnn.model)
    qnn.init_weight_quant_params()
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    qnn.set_quant_state(True, True)
    qnn.set_layer_quant(model=qnn, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")
    g = torch.Generator().manual_seed(31)
    x = torch.randn(2, 4, 64, 64, generator=g).cuda()
    t = torch.tensor([500.0, 500.0], device="cuda")
    y = torch.randn(2, 1, 120, 4096, generator=g).cuda()
    mask = torch.zeros(2, 120, dtype=torch.int64)
    mask[0, :57] = 1
    mask[1, :120] = 1
    mask = mask.cuda()
    with torch.no_grad():
        ref = qnn(x, t, y, mask=mask).float().cpu().numpy()
        n0 = ops.launch_count()
        fused = model.forward_fused(x, t, y, mask=mask).float().cpu().numpy()
    assert np.isfinite(fused).all() and ops.check_status() == 0
    assert ops.launch_count() - n0 >= 2 * 10 + 1      # 10 own launches per block (one per K = 1152 linear) + the patch embedding
    inf, l2 = _rel(fused, ref)
    print("pixart-512 fused (own attention, fused patch embed) vs layerwise schedule: %.3e %.3e" % (inf, l2))
    assert l2 <= 8e-3, (inf, l2)
    # the opt-in INT8 Q/K/V attention on the same model (own tolerance: DESIGN.md 4.2d; default off)
    assert not getattr(model, "attn_int8", False)
    model.attn_int8 = True
    try:
        with torch.no_grad():
            i8 = model.forward_fused(x, t, y, mask=mask).float().cpu().numpy()
    finally:
        model.attn_int8 = False
    inf8, l28 = _rel(i8, fused)
    print("pixart-512 fused with INT8 attention vs fp16 attention: %.3e %.3e" % (inf8, l28))
    assert np.isfinite(i8).all() and 0 < l28 <= 2e-2
