"""viditq_b200.convert.from_reference: the one-call switch from an unmodified reference QuantModel(STDiT) to this repo's
model pair (whose forward_fused is the B200 schedule).  CPU, reference imported from /root/reference (skips on the GPU box);
kernel wrappers swapped for the oracle's stand-ins."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_shims  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="needs /root/reference (not on the GPU box)")

from test_accelerate_cpu import FP_STDIT, stdit_ref, _rel  # noqa: E402,F401


def test_from_reference_carries_weights_checkpoint_and_layer_states(stdit_ref, monkeypatch):  # noqa: F811
    import cpu_ops
    from viditq_b200.convert import from_reference, plain_state_dict
    from viditq_b200.qdiff import QuantLayer
    cpu_ops.patch_ops(monkeypatch)
    build, ckpt, (x, t, y, mask) = stdit_ref
    rq, ref = build()
    # the reference inference flow (quant_txt2video.py:195-207), plus one layer switched to FP and one bit-width changed
    rq.set_quant_params_dict(ckpt)
    rq.set_quant_init_done("weight")
    rq.set_quant_init_done("activation")
    rq.set_quant_state(True, True)
    rq.set_layer_quant(model=rq, module_name_list=FP_STDIT, quant_level="per_layer", weight_quant=False, act_quant=False,
                       prefix="")
    rq.load_bitwidth_config(model=rq, bit_config={"model.blocks.1.mlp.fc2": 6}, bit_type="act")
    rq.half()
    ref.dtype = torch.float16
    rq.cfg_split = True
    keys_before = sorted(rq.state_dict().keys())

    qnn, model = from_reference(rq)
    assert sorted(rq.state_dict().keys()) == keys_before               # the reference objects are untouched
    assert type(model).__name__ == "STDiT" and model.depth == 2 and model.num_spatial == 64 and qnn.cfg_split is True
    # weights: every parameter of the reference, under its original name
    plain = plain_state_dict(ref)
    mine = model.state_dict()
    for k, v in plain.items():
        if k in mine and k.rsplit(".", 1)[-1] in ("weight", "bias", "scale_shift_table", "y_embedding"):
            assert torch.equal(mine[k].float(), v.float()), k
    # checkpoint: every quantiser buffer
    a, b = rq.get_quant_params_dict(), qnn.get_quant_params_dict()
    assert sorted(a.keys()) == sorted(b.keys())
    for name in a:
        for bname, val in a[name][0].items():
            got = b[name][0][bname]
            assert (val is None) == (got is None), (name, bname)
            if val is not None:
                assert torch.equal(val.float(), got.float()), (name, bname)
    # layer states, incl. the FP list, the 6-bit activation layer and init_done
    layers = dict(qnn.quant_layers())
    assert all(isinstance(l, QuantLayer) for l in layers.values()) and len(layers) == 2 * 13 + 6
    assert layers["blocks.0.attn.q"].get_quant_state() == (True, True)
    assert layers["final_layer.linear"].get_quant_state() == (False, False)
    assert layers["blocks.1.mlp.fc2"].act_quantizer.n_bits == 6 and layers["blocks.1.mlp.fc1"].act_quantizer.n_bits == 8
    assert all(l.weight_quantizer.init_done for n, l in layers.items() if n.startswith("blocks."))
    # and it computes what the reference computes: the layer-by-layer forward of the converted model against the
    # reference's own forward (the integer form against the fp16 simulation: inside the 2-block noise band of DESIGN.md 2)
    with torch.no_grad():
        want = rq(x, t, y, mask=mask).float()
        got = qnn(x, t, y, mask=mask).float()
    inf, l2 = _rel(got, want)
    print("from_reference: converted model vs the reference's own forward rel-inf %.3e rel-L2 %.3e" % (inf, l2))
    assert l2 <= 4e-3


from test_accelerate_cpu import pixart_ref  # noqa: E402,F401


def test_from_reference_pixart_with_the_running_stat_layer(pixart_ref, monkeypatch):  # noqa: F811
    """PixArtMS under t2i/configs/quant/alpha/w8a8.yaml: the script state of quant_txt2img.py:291-303 (smooth-quant off except
    the running-stat fc2 of the last block, quirk Q17) survives the conversion, and the converted model tracks the reference
    over consecutive calls (the EMA buffer advances in both)."""
    import cpu_ops
    from viditq_b200.convert import from_reference
    cpu_ops.patch_ops(monkeypatch)
    build, script_state, ckpt, (xs, t, y, mask) = pixart_ref
    rq, ref = build()
    script_state(rq, True, True)
    rq.set_quant_init_done("weight")
    rq.set_quant_init_done("activation")
    rq.set_quant_params_dict(ckpt)
    rq.half()
    qnn, model = from_reference(rq)
    assert type(model).__name__ == "PixArtMS" and model.depth == 2
    layers = dict(qnn.quant_layers())
    last = layers["blocks.1.mlp.fc2"]
    assert last.smooth_quant and last.smooth_quant_running_stat and last.smooth_mode() == "running"
    assert not layers["blocks.0.mlp.fc2"].smooth_quant and layers["final_layer.linear"].get_quant_state() == (True, True)
    for i, x in enumerate(xs):
        with torch.no_grad():
            want = rq(x, t, y, mask=mask).float()
            got = qnn(x, t, y, mask=mask).float()
        inf, l2 = _rel(got, want)
        sa = rq.model.blocks[1].mlp.fc2.act_quantizer.act_scale.float()
        sb = last.act_quantizer.act_scale.float()
        print("from_reference(PixArtMS) call %d: output rel-L2 %.3e, act_scale rel %.3e"
              % (i, l2, ((sa - sb).norm() / sa.norm()).item()))
        assert l2 <= 1.2e-2 and ((sa - sb).norm() / sa.norm()).item() <= 5e-3
