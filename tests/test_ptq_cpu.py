"""PTQ producer (viditq_b200.ptq) against the reference's own PTQ flow: tests/golden/ptq_golden.npz holds
what t2v/scripts/ptq.py:213-362 leaves in ckpt.pth when run with the UNMODIFIED reference classes on a tiny STDiT and a seeded
synthetic calibration set (generator: tests/golden/make_golden_ptq.py; config: w4a8_timestep_aware_cb.yaml's quantiser
sections, two timeranges).  Runs on CPU with the kernel wrappers swapped for the oracle's stand-ins (tests/cpu_ops.py):
what is under test is the producer's host logic — the order of the calibration walk, the EMA statistics of every layer
and rank of input, which quantisers get parameters, the per-timerange weight parameters in fp32 and in the model's fp16."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from test_stdit_graph_cpu import Cfg, FP_LAYERS   # noqa: E402

SMOOTH = dict(alpha=[0.11, 0.31], timerange=[[0, 500], [501, 1000]])


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(ROOT, "tests", "golden", "ptq_golden.npz"))
    return {k: z[k] for k in z.files}


def _calib():
    # the generator's calibration set, re-created from its seed (no reference import: make_golden_ptq.calib_set needs none,
    # but importing the module would import the reference)
    g = torch.Generator().manual_seed(99)
    n = 2 * 2 * 2
    xs = torch.randn(n, 4, 4, 16, 16, generator=g)
    ts = torch.tensor([t for t in (900.0, 100.0) for _ in range(4)])
    cs = torch.randn(n, 1, 120, 4096, generator=g).half().float()
    masks = torch.zeros(n, 120, dtype=torch.int64)
    for i in range(n):
        masks[i, :40 + 9 * i] = 1
    return xs, ts, cs, masks


def _build():
    from viditq_b200.qdiff import QuantModel
    from viditq_b200.stdit import STDiT
    model = STDiT(input_size=(4, 16, 16), depth=2)
    model.init_synthetic(seed=0)
    model.eval()
    sq = Cfg(enable=True, channel_wise_scale_type="momentum_act_max", momentum=0.95, **SMOOTH)
    wq = Cfg(n_bits=4, per_group="channel", channel_dim=0, scale_method="min_max", round_mode="nearest",
             mixed_precision=[4, 6, 8])
    aq = Cfg(n_bits=8, per_group="token", scale_method="min_max", round_mode="nearest_ste", running_stat=False,
             dynamic=True, sym=False, n_spatial_token=model.num_spatial, n_temporal_token=model.num_temporal, n_prompt=120,
             smooth_quant=sq)
    return QuantModel(model, wq, aq), model


def _bufs(ckpt):
    out = {}
    for name, (bufs, params) in ckpt.items():
        assert len(params) == 0
        for bname, val in bufs.items():
            if val is not None:
                out[f"{name}/{bname}"] = val
    return out


def test_run_ptq_reproduces_the_reference_checkpoint(gold, monkeypatch):
    import cpu_ops
    from viditq_b200 import ptq
    cpu_ops.patch_ops(monkeypatch)
    qnn, _ = _build()
    np.random.seed(int(gold["seed"]))
    ckpt = ptq.run_ptq(qnn, _calib(), n_samples=2, batch_size=1, fp_layer_list=FP_LAYERS)
    mine = _bufs(ckpt)
    ref = {k[len("ckpt/"):]: v for k, v in gold.items() if k.startswith("ckpt/")}
    # the same quantisers, and the same buffers filled in (remain_fp layers: statistics only, no weight parameters)
    assert sorted(ckpt.keys()) == sorted(gold["ckpt_names"].tolist())
    assert sorted(mine.keys()) == sorted(ref.keys())
    worst = {"act_scale": 0.0, "delta_list": 0.0, "delta": 0.0}
    flips = total = 0
    for k, r in ref.items():
        m = mine[k].detach().float().numpy()
        assert m.shape == r.shape, (k, m.shape, r.shape)
        kind = k.rsplit("/", 1)[1]
        if kind in ("zero_point_list", "zero_point"):
            flips += int((m != r).sum())
            total += r.size
            assert np.abs(m - r).max() <= 1, k
        else:
            worst[kind] = max(worst[kind], float(np.abs(m - r).max() / np.abs(r).max()))
    print("run_ptq vs reference ckpt: worst relative deviation", worst, "| zero points off by one: %d of %d" % (flips, total))
    # the statistics pass runs this repo's STDiT graph against the reference's: fp32 graphs agree to ~1e-5 (the CPU graph
    # test pins 2e-5 on the output), so do the EMA statistics and everything derived from them
    assert worst["act_scale"] <= 1e-4 and worst["delta_list"] <= 1e-4 and worst["delta"] <= 1e-4
    assert flips <= total * 1e-3
    # model left in the inference state of quant_txt2video.py:195-207
    states = {n: l.get_quant_state() for n, l in qnn.quant_layers()}
    assert states["blocks.0.attn.q"] == (True, True) and states["final_layer.linear"] == (False, False)
    assert all(l.smooth_quant == n.startswith("blocks.") for n, l in qnn.quant_layers())


def test_checkpoint_round_trip_through_the_reference_format(gold, monkeypatch, tmp_path):
    """save_ckpt writes what the reference's torch.save(qnn.get_quant_params_dict()) writes; load_quant_params
    (qdiff/utils.py:65-70) restores every buffer of every quantiser into a fresh QuantModel."""
    import cpu_ops
    from viditq_b200 import ptq
    from viditq_b200.qdiff import load_quant_params
    cpu_ops.patch_ops(monkeypatch)
    qnn, _ = _build()
    np.random.seed(int(gold["seed"]))
    ckpt = ptq.run_ptq(qnn, _calib(), n_samples=2, batch_size=1, fp_layer_list=FP_LAYERS)
    path = str(tmp_path / "ckpt.pth")
    ptq.save_ckpt(ckpt, path)
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert sorted(raw.keys()) == sorted(gold["ckpt_names"].tolist())
    assert all(isinstance(v, (list, tuple)) and len(v) == 2 and len(v[1]) == 0 for v in raw.values())
    qnn2, _ = _build()
    load_quant_params(qnn2, path)
    a, b = _bufs(ckpt), _bufs(qnn2.get_quant_params_dict())
    assert sorted(a) == sorted(b)
    for k in a:
        assert torch.equal(a[k].float(), b[k].float()), k


@pytest.mark.parametrize("dtype", ["fp32", "fp16"])
def test_weight_parameters_bit_exact_given_the_reference_statistics(gold, dtype):
    """Teacher-forced: with the reference's act_scale loaded, the per-timerange weight parameters equal the reference's bit
    for bit — in fp32 (ckpt/...) and in the model's own fp16, which is how ptq.py runs the 16x512x512 config (ckpt16/...)."""
    qnn, model = _build()
    qnn.set_module_name_for_quantizer(module=qnn.model)
    for name, layer in qnn.quant_layers():
        layer.act_quantizer.act_scale = torch.from_numpy(gold[f"ckpt/{name}.act_quantizer/act_scale"])
    if dtype == "fp16":
        qnn.half()
    qnn.set_smooth_quant(True, False)
    qnn.set_layer_smooth_quant(model=qnn, module_name_list=FP_LAYERS, smooth_quant=False, smooth_quant_running_stat=False)
    qnn.set_quant_state(True, False)
    qnn.set_layer_quant(model=qnn, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False, act_quant=False,
                        prefix="")
    qnn.init_weight_quant_params(only_enabled=True, dtype=None)
    prefix = "ckpt/" if dtype == "fp32" else "ckpt16/"
    n = 0
    for name, layer in qnn.quant_layers():
        wq = layer.weight_quantizer
        if not layer.weight_quant:
            assert wq.delta_list is None
            continue
        for b in ("delta_list", "zero_point_list"):
            ref = gold[f"{prefix}{name}.weight_quantizer/{b}"]
            got = getattr(wq, b).detach().numpy()
            assert got.dtype == ref.dtype and np.array_equal(got, ref), (name, b, dtype)
            n += 1
    assert n == 2 * 26


@pytest.mark.parametrize("tag,running", [("static", False), ("static_ema", True)])
def test_static_activation_calibration_against_the_reference_flow(gold, monkeypatch, tag, running):
    """w8a8_naive.yaml's static per-tensor activation quantisers, model in fp16 (how ptq.py runs): the reference walks the
    calibration set with its SIMULATED quantisation, this producer with the integer form (here the oracle's, on CPU) — the
    two agree to <= 1e-3 per layer, so the calibrated ranges agree to that order; zero points to one step."""
    import cpu_ops
    from viditq_b200 import ptq
    from viditq_b200.qdiff import QuantModel
    from viditq_b200.stdit import STDiT
    cpu_ops.patch_ops(monkeypatch)
    model = STDiT(input_size=(4, 16, 16), depth=2)
    model.init_synthetic(seed=0)
    model.eval()
    sq = Cfg(enable=False, channel_wise_scale_type="momentum_act_max", momentum=0.95, alpha=0.625)
    wq = Cfg(n_bits=8, per_group="channel", channel_dim=0, scale_method="min_max", round_mode="nearest")
    aq = Cfg(n_bits=8, per_group=False, scale_method="min_max", round_mode="nearest_ste", running_stat=running,
             dynamic=False, sym=False, n_spatial_token=model.num_spatial, n_temporal_token=model.num_temporal, n_prompt=120,
             smooth_quant=sq)
    qnn = QuantModel(model, wq, aq)
    qnn.half()
    model.dtype = torch.float16
    xs, ts, cs, masks = _calib()
    ckpt = ptq.run_ptq(qnn, (xs, ts, cs.half(), masks), n_samples=2, batch_size=1, fp_layer_list=FP_LAYERS)
    mine = _bufs(ckpt)
    ref = {k[len(tag) + 1:]: v for k, v in gold.items() if k.startswith(tag + "/")}
    assert sorted(mine.keys()) == sorted(ref.keys())
    worst_w = worst_a = 0.0
    zp_off = n_act = 0
    for k, r in ref.items():
        m = mine[k].detach().float().numpy()
        assert m.shape == r.shape, (k, m.shape, r.shape)
        kind = k.rsplit("/", 1)[1]
        if "weight_quantizer" in k:
            assert np.array_equal(m, r), k                      # weight parameters: bit-exact (fp16 arithmetic)
        elif kind in ("delta", "delta_list"):
            worst_a = max(worst_a, float(np.abs(m - r).max() / np.abs(r).max()))
        else:
            zp_off = max(zp_off, float(np.abs(m - r).max()))
            n_act += 1
    print("%s: activation delta worst relative deviation %.2e, zero point worst |diff| %.0f over %d quantisers"
          % (tag, worst_a, zp_off, n_act // 2))
    # last-batch-wins: the range of one tensor that both flows compute to <= 1e-3.  EMA: an fp16 accumulator updated on
    # every calibration forward — a few fp16 ulps (4.9e-4 each) on top of that
    assert worst_a <= (1e-2 if running else 2e-3) and zp_off <= 1
    assert all(l.act_quantizer.init_done and not l.calibrating for _, l in qnn.quant_layers())


def test_timestep_wise_static_calibration_is_refused_not_faked(monkeypatch):
    import cpu_ops
    from viditq_b200 import ptq
    from viditq_b200.qdiff import QuantModel
    from viditq_b200.stdit import STDiT
    cpu_ops.patch_ops(monkeypatch)
    model = STDiT(input_size=(4, 16, 16), depth=1)
    model.init_synthetic(seed=0)
    sq = Cfg(enable=False, channel_wise_scale_type="momentum_act_max", momentum=0.95, alpha=0.625)
    wq = Cfg(n_bits=8, per_group="channel", channel_dim=0, scale_method="min_max", round_mode="nearest")
    aq = Cfg(n_bits=8, per_group=False, scale_method="min_max", round_mode="nearest_ste", running_stat=True,
             dynamic=False, sym=False, n_spatial_token=64, n_temporal_token=4, n_prompt=120, smooth_quant=sq)
    qnn = QuantModel(model, wq, aq)
    qnn.timestep_wise = True
    with pytest.raises(NotImplementedError, match="timestep-wise"):
        ptq.run_ptq(qnn, _calib(), n_samples=2, batch_size=1, fp_layer_list=FP_LAYERS)


def test_reference_quantmodel_loads_the_produced_checkpoint(gold, monkeypatch, tmp_path):
    """Format compatibility in the other direction: the UNMODIFIED reference QuantModel + its own load_quant_params
    (qdiff/utils.py:65-70) read the ckpt.pth this producer wrote and run with it (skips where /root/reference is absent)."""
    from oracle import ref_shims
    if not ref_shims.reference_available():
        pytest.skip("needs /root/reference (not on the GPU box)")
    import cpu_ops
    from viditq_b200 import ptq
    cpu_ops.patch_ops(monkeypatch)
    qnn, mine = _build()
    np.random.seed(int(gold["seed"]))
    path = str(tmp_path / "ckpt.pth")
    ptq.save_ckpt(ptq.run_ptq(qnn, _calib(), n_samples=2, batch_size=1, fp_layer_list=FP_LAYERS), path)
    ref_shims.install_opensora()
    from opensora.models.stdit.stdit import STDiT as RefSTDiT
    from qdiff.models.quant_model import QuantModel as RefQuantModel
    from qdiff.utils import load_quant_params as ref_load
    from viditq_b200.stdit import STDiT
    plain = STDiT(input_size=(4, 16, 16), depth=2)          # un-wrapped: the same seeded weights, reference key names
    plain.init_synthetic(seed=0)
    ref = RefSTDiT(enable_flashattn=False, input_size=(4, 16, 16), depth=2)
    ref.load_state_dict(plain.state_dict(), strict=True)
    ref.eval()
    wq, aq = ref_shims.w8a8_dynamic_configs(n_temporal=ref.num_temporal, n_spatial=ref.num_spatial, n_prompt=120, w_bits=4,
                                            smooth=SMOOTH)
    wq["mixed_precision"] = [4, 6, 8]
    rq = RefQuantModel(ref, wq, aq)
    ref_load(rq, path)
    rq.set_quant_init_done("weight")
    rq.set_quant_init_done("activation")
    rq.set_smooth_quant(smooth_quant=True, smooth_quant_running_stat=False)
    rq.set_layer_smooth_quant(model=rq, module_name_list=FP_LAYERS, smooth_quant=False, smooth_quant_running_stat=False)
    rq.set_quant_state(True, True)
    rq.set_layer_quant(model=rq, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False, act_quant=False,
                       prefix="")
    got = rq.model.blocks[0].attn.q.weight_quantizer.delta_list.float().numpy()
    assert np.allclose(got, gold["ckpt/blocks.0.attn.q.weight_quantizer/delta_list"], rtol=1e-5)
    xs, ts, cs, masks = _calib()
    with torch.no_grad():
        out = rq(xs[:2], ts[:2], cs[:2], mask=masks[:2][::2])
    assert torch.isfinite(out).all()
