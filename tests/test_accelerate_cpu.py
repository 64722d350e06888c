"""The drop-in: viditq_b200.qdiff.accelerate() applied to the UNMODIFIED reference QuantModel (imported from
/root/reference — these tests skip where the reference is absent, i.e. on the GPU box) on CPU, with the kernel wrappers
of viditq_b200.ops swapped for the oracle's integer form (tests/cpu_ops.py).  What is under test is the HOST logic:

  * every reference QuantLayer-family module gets a twin that reads the reference's own quantiser objects, so the
    QuantModel API (set_quant_state, set_layer_quant, load_quant_params, load_bitwidth_config weight AND act,
    set_layer_smooth_quant) keeps working when called AFTER accelerate();
  * per layer, teacher-forced, the accelerated forward equals the reference class's own forward of the same input within
    the north-star tolerance (1e-3 relative, both norms);
  * quirk Q17: the running-stat smooth-quant EMA that t2i/scripts/quant_txt2img.py:297-300 leaves on at inference.
"""
import copy
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_shims  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="needs /root/reference (not on the GPU box)")

FP_STDIT = ["x_embedder", "t_block", "t_embedder", "y_embedder", "final_layer"]
FP_PIXART = ["x_embedder", "t_embedder", "t_block", "y_embedder", "csize_embedder", "ar_embedder"]
TOL = 1e-3      # north star: outputs match the reference's fake-quant path within 1e-3 relative


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max()).item(), ((a - b).norm() / b.norm()).item()


def _teacher_forced(qnn, run, skip=()):
    """Run `run()` with every accelerated reference layer executing BOTH the reference class's own forward and the
    accelerated forward on the same input (the accelerated output flows on).  Returns {layer name: (rel-inf, rel-L2)}."""
    errs, saved = {}, {}

    def make(name, mod, acc):
        def fwd(inp, *a, **k):
            out = acc(inp, *a, **k)
            if mod.weight_quant and mod.act_quant and name not in skip:
                ref = type(mod).forward(mod, inp)            # the unmodified reference code on the reference module
                errs[name] = _rel(out, ref)
            return out
        return fwd
    for name, mod in qnn.model.named_modules():
        if "_viditq_b200" in mod.__dict__:
            saved[mod] = mod.forward
            mod.forward = make(name, mod, saved[mod])
    try:
        with torch.no_grad():
            run()
    finally:
        for mod, f in saved.items():
            mod.forward = f
    return errs


def _worst(errs):
    return max(e[0] for e in errs.values()), max(e[1] for e in errs.values())


@pytest.fixture(scope="module")
def stdit_ref():
    """Reference STDiT (2 blocks) in the reference QuantModel + the quant ckpt of its own PTQ weight pass."""
    ref_shims.install_opensora()
    from opensora.models.stdit.stdit import STDiT as RefSTDiT
    from qdiff.models.quant_model import QuantModel as RefQuantModel
    from viditq_b200.stdit import STDiT
    torch.set_grad_enabled(False)
    cfg = dict(input_size=(4, 16, 16), depth=2)
    mine = STDiT(**cfg)
    mine.init_synthetic(seed=0)
    state = mine.state_dict()

    def build(w_bits=8):
        ref = RefSTDiT(enable_flashattn=False, **cfg)
        ref.load_state_dict(state, strict=True)
        ref.eval()
        wq, aq = ref_shims.w8a8_dynamic_configs(n_temporal=4, n_spatial=64, n_prompt=120, w_bits=w_bits)
        wq["mixed_precision"] = [4, 6, 8]
        qnn = RefQuantModel(ref, wq, aq)
        qnn.set_module_name_for_quantizer(module=qnn.model)
        return qnn, ref
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 4, 4, 16, 16, generator=g)
    y = torch.randn(1, 1, 120, 4096, generator=g).half().float()
    mask = torch.zeros(1, 120, dtype=torch.int64)
    mask[0, :61] = 1
    t = torch.tensor([300.0])
    ckpts = {}
    for w_bits in (8, 4):
        qnn, _ = build(w_bits)
        qnn.set_quant_state(True, False)
        qnn.set_layer_quant(model=qnn, module_name_list=FP_STDIT, quant_level="per_layer", weight_quant=False,
                            act_quant=False, prefix="")
        qnn(x, t, y, mask=mask)                     # PTQ weight pass (ptq.py:266-294)
        qnn.set_quant_init_done("weight")
        qnn.set_quant_init_done("activation")
        ckpts[w_bits] = copy.deepcopy(qnn.get_quant_params_dict())
    build.ckpt4 = ckpts[4]
    return build, ckpts[8], (x, t, y, mask)


def test_accelerate_stdit_state_changes_after_accelerate(stdit_ref, monkeypatch):
    import cpu_ops
    from viditq_b200 import qdiff
    cpu_ops.patch_ops(monkeypatch)
    build, ckpt, (x, t, y, mask) = stdit_ref
    qnn, ref = build()
    n = qdiff.accelerate(qnn)                       # FIRST: everything below happens on the accelerated model
    assert n == 2 * 13 + 6                          # 13 per block + t_embedder(2) + t_block + y_embedder(2) + final_layer
    assert all("_viditq_b200" not in k for k in qnn.state_dict())      # the reference module tree is unchanged
    # inference flow of t2v/scripts/quant_txt2video.py:195-207
    qnn.set_quant_state(True, True)
    qnn.set_layer_quant(model=qnn, module_name_list=FP_STDIT, quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    qnn.set_quant_params_dict(ckpt)                 # == load_quant_params
    qnn.half()
    ref.dtype = torch.float16

    errs = _teacher_forced(qnn, lambda: qnn(x, t, y, mask=mask))
    assert len(errs) == 26
    inf, l2 = _worst(errs)
    print("accelerate(STDiT) per-layer vs reference forward: worst rel-inf %.3e rel-L2 %.3e" % (inf, l2))
    assert inf <= TOL and l2 <= TOL, errs

    # per-layer ACTIVATION widths switched after accelerate (quant_model.py:562-586): dynamic quantiser at 6 / 4 bits
    qnn.load_bitwidth_config(model=qnn, bit_config={"model.blocks.0.mlp.fc2": 6, "model.blocks.1.attn_temp.proj": 4},
                             bit_type="act")
    errs = _teacher_forced(qnn, lambda: qnn(x, t, y, mask=mask))
    inf, l2 = _worst(errs)
    print("after load_bitwidth_config (act 6 / 4 bits): worst rel-inf %.3e rel-L2 %.3e" % (inf, l2))
    assert inf <= TOL and l2 <= TOL, errs

    # a layer switched back to FP after accelerate runs F.linear like the reference
    qnn.set_layer_quant(model=qnn, module_name_list=["blocks.1.mlp"], quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")
    errs2 = _teacher_forced(qnn, lambda: qnn(x, t, y, mask=mask))
    assert "blocks.1.mlp.fc1" not in errs2 and "blocks.1.mlp.fc2" not in errs2 and len(errs2) == 24

    # reloading a different checkpoint after accelerate is picked up (stale prepared weights would not be)
    ck2 = copy.deepcopy(ckpt)
    key = "blocks.0.attn.proj.weight_quantizer"
    ck2[key][0]["delta"] = ck2[key][0]["delta"] * 1.5
    qnn.set_quant_params_dict(ck2)
    qnn.half()
    errs3 = _teacher_forced(qnn, lambda: qnn(x, t, y, mask=mask))
    assert errs3["blocks.0.attn.proj"][1] <= TOL, errs3["blocks.0.attn.proj"]


def test_accelerate_w4_checkpoint_and_weight_bit_switch(stdit_ref, monkeypatch):
    """BASELINE config 4's mechanics on an accelerated reference model: a 4-bit-calibrated checkpoint (W4A8), then
    per-layer WEIGHT widths raised to 8 / 6 bits by load_bitwidth_config after accelerate.  Quirk Q7: delta stays the
    4-bit one, so the '8-bit' layers remain on the 4-bit grid — the twin must re-prepare under the new width and still
    equal the reference."""
    import cpu_ops
    from viditq_b200 import qdiff
    cpu_ops.patch_ops(monkeypatch)
    build, _, (x, t, y, mask) = stdit_ref
    qnn, ref = build(4)
    qdiff.accelerate(qnn)
    qnn.set_quant_state(True, True)
    qnn.set_layer_quant(model=qnn, module_name_list=FP_STDIT, quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    qnn.set_quant_params_dict(build.ckpt4)
    qnn.half()
    ref.dtype = torch.float16
    errs = _teacher_forced(qnn, lambda: qnn(x, t, y, mask=mask))
    inf, l2 = _worst(errs)
    print("accelerate(STDiT) W4A8 per-layer: worst rel-inf %.3e rel-L2 %.3e" % (inf, l2))
    assert len(errs) == 26 and inf <= TOL and l2 <= TOL, errs
    tw = dict(qnn.model.named_modules())["blocks.0.mlp.fc1"]._viditq_b200
    codes4 = tw._prepared[(4, 0)][1].codes.clone()
    qnn.load_bitwidth_config(model=qnn, bit_config={"model.blocks.0.mlp.fc1": 8, "model.blocks.1.attn.q": 6},
                             bit_type="weight")
    errs = _teacher_forced(qnn, lambda: qnn(x, t, y, mask=mask))
    inf, l2 = _worst(errs)
    print("after load_bitwidth_config (weight 8 / 6 bits on the 4-bit grid): worst rel-inf %.3e rel-L2 %.3e" % (inf, l2))
    assert inf <= TOL and l2 <= TOL, errs
    assert tw.weight_quantizer.n_bits == 8 and (8, 0) in tw._prepared
    # Q7: still the 4-bit grid — the only codes that may move are range ends whose two roundings add up to 16 (> 15)
    c8 = tw._prepared[(8, 0)][1].codes
    assert int(c8.max()) <= 16 and float((c8 != codes4).float().mean()) < 1e-3


def test_accelerate_matches_reference_end_to_end(stdit_ref, monkeypatch):
    """Whole forward: accelerated model vs a second, untouched reference model.  End to end the distance sits in the
    re-quantisation noise band; stated next to the quantisation error itself."""
    import cpu_ops
    from viditq_b200 import qdiff
    cpu_ops.patch_ops(monkeypatch)
    build, ckpt, (x, t, y, mask) = stdit_ref
    outs = []
    for accel in (False, True):
        qnn, ref = build()
        if accel:
            qdiff.accelerate(qnn)
        qnn.set_quant_state(True, True)
        qnn.set_layer_quant(model=qnn, module_name_list=FP_STDIT, quant_level="per_layer", weight_quant=False,
                            act_quant=False, prefix="")
        qnn.set_quant_init_done("weight")
        qnn.set_quant_init_done("activation")
        qnn.set_quant_params_dict(ckpt)
        qnn.half()
        ref.dtype = torch.float16
        outs.append(qnn(x, t, y, mask=mask))
        if not accel:
            qnn.set_quant_state(False, False)
            fp = qnn(x, t, y, mask=mask)
    d = _rel(outs[1], outs[0])
    q = _rel(outs[0], fp)
    print("accelerated vs reference W8A8 forward: rel-inf %.3e rel-L2 %.3e (quantisation error itself %.3e)" % (d + (q[1],)))
    assert d[1] <= 4e-3 and d[1] < 0.6 * q[1], (d, q)


@pytest.fixture(scope="module")
def pixart_ref():
    ref_shims.install_pixart()
    from diffusion.model.nets.PixArtMS import PixArtMS as RefPixArt
    from qdiff.models.quant_model import QuantModel as RefQuantModel
    from viditq_b200.pixart import PixArtMS
    torch.set_grad_enabled(False)
    cfg = dict(input_size=16, depth=2)
    mine = PixArtMS(**cfg)
    mine.init_synthetic(seed=0)
    state = mine.state_dict()
    SMOOTH_LAYERS = ["blocks.1.mlp.fc2"]            # the role of blocks.27.mlp.fc2 in the 28-block model

    def build():
        ref = RefPixArt(**cfg)
        ref.load_state_dict(state, strict=True)
        ref.eval()
        # t2i/configs/quant/alpha/w8a8.yaml: dynamic per-token W8A8, smooth-quant momentum_act_max / 0.95 / alpha 0.3
        wq, aq = ref_shims.w8a8_dynamic_configs(n_temporal=1, n_spatial=64, n_prompt=120, smooth=dict(alpha=0.3))
        qnn = RefQuantModel(ref, wq, aq, model_type="pixart")
        qnn.set_module_name_for_quantizer(module=qnn.model)
        return qnn, ref

    def script_state(qnn, wq_on, aq_on):
        """t2i/scripts/quant_txt2img.py:291-303 (and ptq.py:222-227)."""
        qnn.set_quant_state(wq_on, aq_on)
        qnn.set_layer_quant(model=qnn, module_name_list=FP_PIXART, quant_level="per_layer", weight_quant=False,
                            act_quant=False, prefix="")
        qnn.set_smooth_quant(smooth_quant=False, smooth_quant_running_stat=False)
        qnn.set_layer_smooth_quant(model=qnn, module_name_list=SMOOTH_LAYERS, smooth_quant=True,
                                   smooth_quant_running_stat=True)
    g = torch.Generator().manual_seed(23)
    z = torch.randn(1, 4, 16, 16, generator=g)
    xs = [torch.cat([z, z], 0) * s for s in (1.0, 0.7, 1.3)]
    y = torch.randn(2, 1, 120, 4096, generator=g).half().float()
    mask = torch.zeros(2, 120, dtype=torch.int64)
    mask[:, :77] = 1
    t = torch.tensor([500.0, 500.0])
    qnn, _ = build()
    script_state(qnn, True, False)
    qnn(xs[0], t, y, mask=mask)                     # PTQ weight pass: also collects act_scale for the smooth-quant layer
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    ckpt = copy.deepcopy(qnn.get_quant_params_dict())
    assert ckpt["blocks.1.mlp.fc2.act_quantizer"][0]["act_scale"] is not None
    return build, script_state, ckpt, (xs, t, y, mask)


def test_accelerate_pixart_w8a8_yaml_with_running_stat_smooth_quant(pixart_ref, monkeypatch):
    """BASELINE config 2 under the ViDiT-Q W8A8 config: the script flow of quant_txt2img.py runs through accelerate(),
    including the stateful layer (quirk Q17).  Two models (reference / accelerated) see the same three consecutive
    inputs; outputs stay inside the re-quantisation band and the EMA buffers track each other."""
    import cpu_ops
    from viditq_b200 import qdiff
    cpu_ops.patch_ops(monkeypatch)
    build, script_state, ckpt, (xs, t, y, mask) = pixart_ref
    models = []
    for accel in (False, True):
        qnn, ref = build()
        if accel:
            assert qdiff.accelerate(qnn) == 2 * 7 + 6          # 7 per block; t_embedder (2), t_block, y_embedder (2), final_layer; the conv x_embedder stays
        script_state(qnn, True, True)
        qnn.set_quant_init_done("weight")
        qnn.set_quant_init_done("activation")
        qnn.set_quant_params_dict(ckpt)
        qnn.half()
        models.append(qnn)
    ref_q, acc_q = models
    # the reference's OWN noise band on this model: its fp16 simulation against the same simulation with identical
    # codes but un-rounded dequantised operands (oracle.torch_fake_quant, exact=True), smooth-quant off in both
    band_q, _ = build()
    script_state(band_q, True, True)
    band_q.set_smooth_quant(False, False)
    band_q.set_quant_init_done("weight")
    band_q.set_quant_init_done("activation")
    band_q.set_quant_params_dict(ckpt)
    band_q.half()
    o_sim = band_q(xs[0], t, y, mask=mask)
    band_q.set_quant_state(False, False)
    o_fp = band_q(xs[0], t, y, mask=mask)
    script_state(band_q, True, True)
    band_q.set_smooth_quant(False, False)
    from oracle import torch_fake_quant as TF
    for mod in band_q.model.modules():
        if hasattr(mod, "weight_quantizer") and isinstance(mod.org_module, torch.nn.Linear):
            def exact(inp, _m=mod, **k):
                if not (_m.weight_quant and _m.act_quant):
                    return type(_m).forward(_m, inp)
                wq_ = _m.weight_quantizer
                return TF.quant_linear_fake(inp.float(), _m.weight.float(), _m.bias.float(), wq_.delta.float(),
                                            wq_.zero_point.float(), wq_.n_bits, 8, exact=True).half()
            mod.forward = exact
    band = _rel(band_q(xs[0], t, y, mask=mask).float(), o_sim.float())
    qerr = _rel(o_sim.float(), o_fp.float())
    print("pixart reference noise band (fp16 sim vs exact-operand sim) rel-L2 %.3e; quantisation error %.3e" % (band[1], qerr[1]))
    fc2_r = dict(ref_q.model.named_modules())["blocks.1.mlp.fc2"]
    fc2_a = dict(acc_q.model.named_modules())["blocks.1.mlp.fc2"]
    assert fc2_a.smooth_quant and fc2_a.smooth_quant_running_stat
    s0 = fc2_r.act_quantizer.act_scale.clone()
    for i, x in enumerate(xs):
        o_r = ref_q(x, t, y, mask=mask)
        o_a = acc_q(x, t, y, mask=mask)
        d = _rel(o_a.float(), o_r.float())
        sd = _rel(fc2_a.act_quantizer.act_scale.float(), fc2_r.act_quantizer.act_scale.float())
        print("pixart call %d: accelerated vs reference rel-inf %.3e rel-L2 %.3e | act_scale EMA rel-inf %.3e" % (i, d[0], d[1], sd[0]))
        assert d[1] <= band[1], (d, band)   # inside the reference's own fp16-simulation noise band (measured above)
        assert sd[0] <= 2e-2, sd            # the EMA sees slightly different inputs (upstream re-quantisation noise)
    assert not torch.equal(fc2_r.act_quantizer.act_scale, s0)      # the state really advanced
    # teacher-forced per-layer parity on the accelerated model (the stateful layer is compared separately below)
    errs = _teacher_forced(acc_q, lambda: acc_q(xs[0], t, y, mask=mask), skip=("blocks.1.mlp.fc2",))
    inf, l2 = _worst(errs)
    print("accelerate(PixArt) per-layer (%d layers): worst rel-inf %.3e rel-L2 %.3e" % (len(errs), inf, l2))
    assert len(errs) == 14 and inf <= TOL and l2 <= TOL, errs


def test_running_stat_layer_tracks_the_reference_bit_for_bit(pixart_ref, monkeypatch):
    """Q17 at layer level: the same fp16 inputs through the reference layer and through its accelerated copy, three calls:
    identical act_scale buffers after every call (same EMA arithmetic on exact column maxima) and outputs <= 1e-3."""
    import cpu_ops
    from viditq_b200 import qdiff
    cpu_ops.patch_ops(monkeypatch)
    build, script_state, ckpt, _ = pixart_ref
    pair = []
    for accel in (False, True):
        qnn, _ = build()
        if accel:
            qdiff.accelerate(qnn)
        script_state(qnn, True, True)
        qnn.set_quant_init_done("weight")
        qnn.set_quant_init_done("activation")
        qnn.set_quant_params_dict(ckpt)
        qnn.half()
        qnn.set_timestep_id_for_quantlayer(500.0) if hasattr(qnn, "set_timestep_id_for_quantlayer") else None
        pair.append(dict(qnn.model.named_modules())["blocks.1.mlp.fc2"])
    ref_l, acc_l = pair
    g = torch.Generator().manual_seed(5)
    for i in range(3):
        x = (torch.randn(2, 64, 4608, generator=g) * (1 + i)).half()
        x[..., 7] *= 9
        o_r, o_a = ref_l(x), acc_l(x)
        assert torch.equal(ref_l.act_quantizer.act_scale, acc_l.act_quantizer.act_scale), i
        d = _rel(o_a, o_r)
        print("running-stat fc2 call %d: rel-inf %.3e rel-L2 %.3e" % (i, d[0], d[1]))
        assert d[0] <= TOL and d[1] <= TOL, d


def test_dynamic_channel_scale_type(stdit_ref, monkeypatch):
    """channel_wise_scale_type 'dynamic' (quant_layer.py:115-116): the channel scale comes from the live input."""
    import cpu_ops
    from viditq_b200 import qdiff
    cpu_ops.patch_ops(monkeypatch)
    ref_shims.install_opensora()
    from qdiff.models.quant_layer import QuantLayer as RefQuantLayer
    torch.manual_seed(3)
    lin = torch.nn.Linear(1152, 384)
    wq, aq = ref_shims.w8a8_dynamic_configs(smooth=dict(alpha=0.4))
    aq["smooth_quant"]["channel_wise_scale_type"] = "dynamic"
    layers = []
    for accel in (False, True):
        l = RefQuantLayer(copy.deepcopy(lin), wq, aq)
        l.cur_timestep_id = 10
        layers.append(l)
    ref_l, acc_src = layers

    class Holder(torch.nn.Module):
        def __init__(self, l):
            super().__init__()
            self.model = torch.nn.Sequential(l)
    x = torch.randn(2, 96, 1152)
    x[..., 100] *= 15
    for l in layers:       # weight init pass in fp32 with the smooth scale of this very input (reference logic)
        l.set_quant_state(True, False)
        l(x)
        l.weight_quantizer.init_done = True
        l.act_quantizer.init_done = True
        l.set_quant_state(True, True)
        l.half()
    assert qdiff.accelerate(Holder(acc_src)) == 1
    xh = x.half()
    d = _rel(acc_src(xh), ref_l(xh))
    print("dynamic smooth-quant scale: rel-inf %.3e rel-L2 %.3e" % d)
    assert d[0] <= TOL and d[1] <= TOL, d
