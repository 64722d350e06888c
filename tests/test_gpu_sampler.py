"""viditq_b200.sampler.GraphedSampler: the complete DDIM loop on the fused schedule with one CUDA-graph replay per step,
against the eager loop (ddim_sample_loop) step for step — bit-identical latents — and one captured graph per
(smooth-quant timerange, mixed-precision range) key."""
import functools
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from test_stdit_graph_cpu import Cfg, FP_LAYERS   # noqa: E402


def _model(smooth):
    from viditq_b200.qdiff import QuantModel
    from viditq_b200.stdit import STDiT
    T, S = 4, 256
    model = STDiT(input_size=(T, 32, 32), depth=2)
    model.init_synthetic(seed=3)
    model.eval()
    sq = Cfg(enable=False, channel_wise_scale_type="momentum_act_max", momentum=0.95, alpha=0.625)
    if smooth:
        sq = Cfg(enable=True, channel_wise_scale_type="momentum_act_max", momentum=0.95, alpha=[0.11, 0.31],
                 timerange=[[0, 500], [501, 1000]])
    wq = Cfg(n_bits=4 if smooth else 8, per_group="channel", channel_dim=0, scale_method="min_max", round_mode="nearest",
             mixed_precision=[4, 6, 8])
    aq = Cfg(n_bits=8, per_group="token", scale_method="min_max", round_mode="nearest_ste", running_stat=False, dynamic=True,
             sym=False, n_spatial_token=S, n_temporal_token=T, n_prompt=120, smooth_quant=sq)
    qnn = QuantModel(model, wq, aq)
    if smooth:
        g = torch.Generator().manual_seed(5)
        for name, layer in qnn.quant_layers():
            layer.act_quantizer.act_scale = torch.rand(2, 1, layer.in_features, generator=g) + 0.5
            if not name.startswith("blocks."):
                layer.smooth_quant = False
    qnn.cuda()
    qnn.half()
    model.dtype = torch.float16
    qnn.set_module_name_for_quantizer(module=qnn.model)
    qnn.fp_layer_list = FP_LAYERS
    qnn.init_weight_quant_params()
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    qnn.set_quant_state(True, True)
    return qnn, model


@pytest.mark.gpu
@pytest.mark.parametrize("smooth", [False, True])
def test_graphed_sampling_loop_equals_the_eager_loop(smooth):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from viditq_b200 import ops
    from viditq_b200.sampler import GraphedSampler, SpacedDDIM, ddim_sample_loop
    qnn, model = _model(smooth)
    ddim = SpacedDDIM(num_sampling_steps=6, cfg_scale=4.0)        # timesteps 999 ... 0: both timeranges of the W4A8 config
    g = torch.Generator().manual_seed(8)
    z0 = torch.randn(1, 4, 4, 32, 32, generator=g).cuda()
    yc = torch.randn(1, 1, 120, 4096, generator=g).cuda()
    yu = torch.randn(1, 1, 120, 4096, generator=g).cuda()
    mask = torch.zeros(1, 120, dtype=torch.int64)
    mask[0, :93] = 1
    mask = mask.cuda()
    eager = []
    with torch.no_grad():
        want = ddim_sample_loop(ddim, model.forward_fused, z0.clone(), yc, yu, mask, qnn=qnn,
                                on_step=lambda i, z: eager.append(z.clone()),
                                stacked_forward=functools.partial(model.forward_fused, independent=True))
    gs = GraphedSampler(qnn, model, ddim, yc, yu, mask, z0.shape)
    graphed = []
    got = gs.sample(z0, on_step=lambda i, z: graphed.append(z.clone()))
    assert ops.check_status() == 0 and torch.isfinite(got).all()
    assert len(gs.graphs) == (2 if smooth else 1) and len(graphed) == ddim.num_timesteps == 6
    for a, b in zip(graphed, eager):
        assert torch.equal(a, b)
    assert torch.equal(got, want)
    again = gs.sample(z0)                                           # the captured graphs are reusable
    assert torch.equal(again, want) and len(gs.graphs) == (2 if smooth else 1)


@pytest.mark.gpu
def test_graphed_sampling_loop_with_per_timestep_mixed_precision():
    """Config 4's mechanics inside the graphed loop: the per-timestep bit tables (quant_txt2video_mp.py:533-540) switch layer
    widths when the step index enters a new range — a new key, hence a new captured graph; latents equal the eager loop's."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from viditq_b200.sampler import GraphedSampler, SpacedDDIM, ddim_sample_loop
    qnn, model = _model(True)
    names = ["model." + n for n, _ in qnn.quant_layers() if n.startswith("blocks.")]
    tab = {n: (8 if ".mlp." in n else 4) for n in names}
    qnn.timestep_wise_mp = True
    qnn.time_mp_config_weight = {"5-4": {n: 8 for n in names}, "3-0": tab, "fp_layers": {"5-4": FP_LAYERS, "3-0": FP_LAYERS}}
    qnn.time_mp_config_act = {"5-4": {n: 8 for n in names}, "3-0": {n: 8 for n in names}}
    ddim = SpacedDDIM(num_sampling_steps=6, cfg_scale=7.0)
    g = torch.Generator().manual_seed(9)
    z0 = torch.randn(1, 4, 4, 32, 32, generator=g).cuda()
    yc = torch.randn(1, 1, 120, 4096, generator=g).cuda()
    yu = torch.randn(1, 1, 120, 4096, generator=g).cuda()
    mask = torch.zeros(1, 120, dtype=torch.int64)
    mask[0, :77] = 1
    mask = mask.cuda()
    with torch.no_grad():
        want = ddim_sample_loop(ddim, model.forward_fused, z0.clone(), yc, yu, mask, qnn=qnn,
                                stacked_forward=functools.partial(model.forward_fused, independent=True))
    gs = GraphedSampler(qnn, model, ddim, yc, yu, mask, z0.shape)
    got = gs.sample(z0)
    # timesteps 999, 799 (range 5-4, timerange 1) | 599 (3-0, timerange 1) | 400, 200, 0 (3-0, timerange 0): three keys
    assert sorted(gs.graphs.keys()) == [(0, "3-0"), (1, "3-0"), (1, "5-4")]
    assert torch.equal(got, want)
