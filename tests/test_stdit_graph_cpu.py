"""CPU tests of the host-side graph (viditq_b200.stdit) and operator-API mirror (viditq_b200.qdiff): no GPU needed
because un-quantised QuantLayers run F.linear.  Pins the graph (embedders, pos-emb, (T S) layouts, cross-attention
segments, unpatchify) against the reference STDiT's fp32 output stored in tests/golden/stdit_small_golden.npz."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP_LAYERS = ["x_embedder", "t_block", "t_embedder", "y_embedder", "final_layer"]


class Cfg(dict):
    __getattr__ = dict.get


def quant_cfgs(T, S):
    sq = Cfg(enable=False, channel_wise_scale_type="momentum_act_max", momentum=0.95, alpha=0.625)
    wq = Cfg(n_bits=8, per_group="channel", channel_dim=0, scale_method="min_max", round_mode="nearest",
             mixed_precision=[4, 6, 8])
    aq = Cfg(n_bits=8, per_group="token", scale_method="min_max", round_mode="nearest_ste", running_stat=False,
             dynamic=True, sym=False, n_spatial_token=S, n_temporal_token=T, n_prompt=120, smooth_quant=sq)
    return wq, aq


@pytest.fixture(scope="module")
def small():
    z = np.load(os.path.join(ROOT, "tests", "golden", "stdit_small_golden.npz"))
    return {k: z[k] for k in z.files}


def ckpt_from_golden(g):
    ck = {str(n): [OrderedDict(), OrderedDict()] for n in g["ckpt_names"]}
    for key, val in g.items():
        if key.startswith("ckpt/"):
            _, name, buf = key.split("/")
            ck.setdefault(name, [OrderedDict(), OrderedDict()])[0][buf] = torch.from_numpy(val)
    for name in ck:
        for buf in ("delta_list", "zero_point_list", "delta", "zero_point", "alpha"):
            ck[name][0].setdefault(buf, None)
    return ck


def build_qnn(g):
    from viditq_b200.qdiff import QuantModel
    from viditq_b200.stdit import STDiT
    model = STDiT(input_size=(4, 16, 16), depth=2)
    model.init_synthetic(seed=0)
    model.eval()
    wq, aq = quant_cfgs(int(g["T"]), int(g["S"]))
    qnn = QuantModel(model, wq, aq)
    return qnn, model


def test_layer_replacement_rules_match_reference():
    from viditq_b200 import qdiff
    g = {"T": 4, "S": 64}
    qnn, model = build_qnn(g)
    b = model.blocks[0]
    assert type(b.attn.q) is qdiff.QuantSpatialAttnLinear and type(b.attn.proj) is qdiff.QuantSpatialAttnLinear
    assert type(b.attn_temp.k) is qdiff.QuantTemporalAttnLinear
    assert type(b.cross_attn.kv_linear) is qdiff.QuantCrossAttnLinear
    assert type(b.mlp.fc1) is qdiff.QuantLayer and type(model.final_layer.linear) is qdiff.QuantLayer
    n_block_layers = sum(1 for n, _ in qnn.quant_layers() if n.startswith("blocks."))
    assert n_block_layers == 13 * 2     # 13 quantised linears per STDiT block


def test_fp_graph_matches_reference_fp32(small):
    qnn, model = build_qnn(small)
    qnn.set_quant_state(False, False)
    with torch.no_grad():
        out = qnn(torch.from_numpy(small["x"]), torch.from_numpy(small["t"]),
                  torch.from_numpy(small["y"]).float(), mask=torch.from_numpy(small["mask"]))
    ref = small["out_fp32"]
    assert out.shape == ref.shape
    err = np.abs(out.numpy() - ref).max() / np.abs(ref).max()
    assert err < 2e-5, err


def test_reference_ckpt_loads_unchanged_and_state_api(small):
    from viditq_b200 import qdiff
    qnn, model = build_qnn(small)
    qnn.set_module_name_for_quantizer(module=qnn.model)
    qnn.set_quant_params_dict(ckpt_from_golden(small))
    wq = model.blocks[1].mlp.fc2.weight_quantizer
    assert wq.delta.shape == (1152, 1) and wq.delta_list.shape == (3, 1, 1152, 1)
    assert wq.module_name == "blocks.1.mlp.fc2.weight_quantizer"
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    assert wq.init_done and model.blocks[0].attn.q.act_quantizer.init_done
    qnn.set_quant_state(True, True)
    qnn.set_layer_quant(model=qnn, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")
    assert model.blocks[0].attn.q.get_quant_state() == (True, True)
    assert model.final_layer.linear.get_quant_state() == (False, False)
    assert model.y_embedder.y_proj.fc1.get_quant_state() == (False, False)
    # round trip of the ckpt dict format
    out = qnn.get_quant_params_dict()
    assert set(out) == set(ckpt_from_golden(small))
    # bit-width switch keeps delta (quirk Q7)
    d0 = wq.delta.clone()
    qnn.load_bitwidth_config(qnn.model, {"blocks.1.mlp.fc2": 4}, "weight")
    assert wq.n_bits == 4 and wq.bit_idx == 0 and torch.equal(wq.delta, d0)
    # quantised forward without a GPU must fail loudly, never fall back
    with pytest.raises(Exception):
        model.blocks[0].mlp.fc1(torch.zeros(1, 4, 1152))
    assert isinstance(model.blocks[0].attn.q.act_quantizer, qdiff.DynamicActQuantizer)


def test_min_max_weight_init_matches_reference_ckpt(small):
    """QuantModel.init_weight_quant_params restates the PTQ weight pass: same delta / zero-point as the reference."""
    qnn, model = build_qnn(small)
    qnn.set_module_name_for_quantizer(module=qnn.model)
    qnn.init_weight_quant_params()
    ck = ckpt_from_golden(small)
    for name in ["blocks.0.attn.q", "blocks.1.cross_attn.kv_linear", "blocks.1.mlp.fc2"]:
        mod = dict(qnn.quant_layers())[name]
        ref = ck[name + ".weight_quantizer"][0]
        np.testing.assert_allclose(mod.weight_quantizer.delta_list.numpy(), ref["delta_list"].numpy(), rtol=1e-6)
        np.testing.assert_array_equal(mod.weight_quantizer.zero_point_list.numpy(), ref["zero_point_list"].numpy())


def test_pattern_in_semantics():
    from viditq_b200.qdiff import pattern_in
    assert pattern_in("model.blocks.3.attn.q", "blocks.[0-5].attn")
    assert not pattern_in("model.blocks.7.attn.q", "blocks.[0-5].attn.q")
    assert pattern_in("blocks.0.mlp.fc1", "*.0.mlp")
    assert pattern_in("x_embedder", "x_embedder")


def test_frame_sharding_host_helpers():
    """Host-side pieces of the frame-sharded forward that need no process group: the frame ranges, the unpatchify of a
    rank's frames, and the refusal of batch-pooled codes."""
    from viditq_b200 import shard
    from viditq_b200.stdit import STDiT
    assert [shard.frame_slice(16, 4, r) for r in range(4)] == [(0, 4), (4, 8), (8, 12), (12, 16)]
    with pytest.raises(ValueError):
        shard.frame_slice(16, 3, 0)
    model = STDiT(input_size=(4, 8, 8), depth=1, hidden_size=64, num_heads=4)
    full = torch.randn(2, 4 * 16, 4 * 8)                       # [B, T*S, patch * out_channels]
    whole = model.unpatchify(full)
    part = model.unpatchify(full.view(2, 4, 16, 32)[:, 1:3].reshape(2, 2 * 16, 32))   # frames 1..2 only
    assert part.shape == (2, 8, 2, 8, 8) and torch.equal(part, whole[:, :, 1:3])

    class A:
        G = 2
    with pytest.raises(ValueError):
        shard.exchange_act_codes(A(), 1, 2, 16, 2, True)


def test_forward_fused_refuses_states_it_does_not_implement(small):
    """the fused schedule hard-wires dynamic per-token W+A quantisation of all 13 block linears.  A
    static-scale checkpoint (w8a8_naive.yaml), a layer switched to FP, or q/k/v with different activation widths must raise
    on EVERY call (before any kernel runs: this test needs no GPU) instead of silently mixing quantiser types."""
    from viditq_b200.qdiff import QuantModel
    from viditq_b200.stdit import STDiT
    x, t = torch.from_numpy(small["x"]), torch.from_numpy(small["t"])
    y, mask = torch.from_numpy(small["y"]).float(), torch.from_numpy(small["mask"])

    def fresh(dynamic=True, per_group="token"):
        model = STDiT(input_size=(4, 16, 16), depth=2)
        model.eval()
        wq, aq = quant_cfgs(int(small["T"]), int(small["S"]))
        aq = Cfg(aq, dynamic=dynamic, per_group=per_group)
        qnn = QuantModel(model, wq, aq)
        qnn.set_quant_state(True, True)
        qnn.set_layer_quant(model=qnn, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                            act_quant=False, prefix="")
        return qnn, model
    qnn, model = fresh(dynamic=False, per_group=False)               # static per-tensor activation scales
    with torch.no_grad(), pytest.raises(NotImplementedError, match="static"):
        model.forward_fused(x, t, y, mask=mask)
    qnn, model = fresh(dynamic=False, per_group="token")             # static per-token: MASK_SELECT=False layouts
    with torch.no_grad(), pytest.raises(NotImplementedError, match="static per-token"):
        model.forward_fused(x, t, y, mask=mask)
    with torch.no_grad(), pytest.raises(NotImplementedError, match="static per-token"):
        model.forward_fused(x, t, y, mask=mask, plan=model.mask_select_plan(mask))     # the guard does not depend on `plan`
    qnn, model = fresh()
    qnn.set_layer_quant(model=qnn, module_name_list=["blocks.1.cross_attn.proj"], quant_level="per_layer",
                        weight_quant=False, act_quant=False, prefix="")
    with torch.no_grad(), pytest.raises(NotImplementedError, match="blocks.1.cross_attn.proj"):
        model.forward_fused(x, t, y, mask=mask)
    qnn, model = fresh()
    qnn.load_bitwidth_config(qnn.model, {"blocks.0.attn_temp.k": 6}, "act")
    with torch.no_grad(), pytest.raises(NotImplementedError, match="activation widths differ"):
        model.forward_fused(x, t, y, mask=mask)


def test_mask_select_false_branch_matches_reference_layout(small):
    """Quirk Q14 (stdit.py:272-300): with a static per-token activation quantiser the prompt rows are zero-masked, not
    dropped — every sample keeps max_len rows so that calibrated per-token (delta, zp) keep their positions."""
    from viditq_b200.qdiff import QuantModel
    from viditq_b200.stdit import STDiT
    model = STDiT(input_size=(4, 16, 16), depth=1)
    model.eval()
    wq, aq = quant_cfgs(int(small["T"]), int(small["S"]))
    QuantModel(model, wq, Cfg(aq, dynamic=False, per_group="token"))
    x, t = torch.from_numpy(small["x"]), torch.from_numpy(small["t"])
    y, mask = torch.from_numpy(small["y"]).float(), torch.from_numpy(small["mask"])
    with torch.no_grad():
        _, _, _, yy, y_lens = model.embed(x, t, y, mask)
        full = model.y_embedder(y).squeeze(1)
    n_keep = int(mask.sum())
    assert y_lens == [120] and yy.shape == (1, 120, 1152)
    assert torch.equal(yy[0, :n_keep], full[0, :n_keep]) and float(yy[0, n_keep:].abs().max()) == 0.0
