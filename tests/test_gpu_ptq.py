"""PTQ producer on the GPU (viditq_b200.ptq; row N4): the calibration walk in fp16 on the device — statistics through the
vq_col_absmax kernel — against the reference's PTQ result (tests/golden/ptq_golden.npz, an fp32 CPU run of the unmodified
reference: the comparison carries the fp16-vs-fp32 difference of the statistics pass, ~1e-3), then the checkpoint round
trip: save -> load into a fresh QuantModel -> identical denoiser output on the fused schedule."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from test_ptq_cpu import _build, _calib, _bufs     # noqa: E402
from test_stdit_graph_cpu import FP_LAYERS          # noqa: E402


@pytest.mark.gpu
def test_ptq_on_the_device_and_checkpoint_round_trip(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from viditq_b200 import ops, ptq
    from viditq_b200.qdiff import load_quant_params
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ptq_golden.npz"))
    qnn, model = _build()
    qnn.cuda()
    qnn.half()
    model.dtype = torch.float16
    xs, ts, cs, masks = _calib()
    np.random.seed(int(gold["seed"]))
    n0 = ops.launch_count()
    ckpt = ptq.run_ptq(qnn, (xs, ts, cs.half(), masks), n_samples=2, batch_size=1, fp_layer_list=FP_LAYERS,
                       device=torch.device("cuda"))
    assert ops.launch_count() - n0 >= 4 * 32          # >= one vq_col_absmax per QuantLayer per calibration forward
    mine = _bufs(ckpt)
    worst = {"act_scale": 0.0, "delta_list": 0.0}
    for k in gold.files:
        if not k.startswith("ckpt/"):
            continue
        kind = k.rsplit("/", 1)[1]
        if kind in worst:
            m, r = mine[k[len("ckpt/"):]].detach().float().cpu().numpy(), gold[k]
            worst[kind] = max(worst[kind], float(np.linalg.norm(m - r) / np.linalg.norm(r)))
    print("PTQ on the device (fp16) vs the reference's fp32 CPU run: rel-L2 act_scale %.2e, weight delta %.2e"
          % (worst["act_scale"], worst["delta_list"]))
    assert worst["act_scale"] <= 5e-3 and worst["delta_list"] <= 5e-3
    # ---- round trip through ckpt.pth
    path = str(tmp_path / "ckpt.pth")
    ptq.save_ckpt(ckpt, path)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 4, 4, 16, 16, generator=g).cuda()
    y = torch.randn(1, 1, 120, 4096, generator=g).cuda()
    mask = torch.zeros(1, 120, dtype=torch.int64)
    mask[0, :61] = 1
    mask = mask.cuda()
    t = torch.tensor([700.0], device="cuda")
    with torch.no_grad():
        qnn.set_timestep_id_for_quantlayer(700.0)
        a = model.forward_fused(x, t, y, mask=mask).float().cpu()
    qnn2, model2 = _build()
    load_quant_params(qnn2, path, dtype=torch.float16)      # qdiff/utils.py:65-70 loads on the CPU; the model moves after
    qnn2.cuda()
    qnn2.half()
    model2.dtype = torch.float16
    qnn2.set_quant_init_done("weight")
    qnn2.set_quant_init_done("activation")
    qnn2.set_smooth_quant(True, False)
    qnn2.set_layer_smooth_quant(model=qnn2, module_name_list=FP_LAYERS, smooth_quant=False, smooth_quant_running_stat=False)
    qnn2.set_quant_state(True, True)
    qnn2.set_layer_quant(model=qnn2, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                         act_quant=False, prefix="")
    with torch.no_grad():
        qnn2.set_timestep_id_for_quantlayer(700.0)
        b = model2.forward_fused(x, t, y, mask=mask).float().cpu()
    assert torch.isfinite(a).all() and ops.check_status() == 0
    assert torch.equal(a, b)


@pytest.mark.gpu
def test_static_checkpoint_from_the_producer_runs_on_the_fused_schedule():
    """w8a8_naive.yaml family end to end on the device: static per-tensor activation scales calibrated by
    viditq_b200.ptq.run_ptq (through the integer kernels), then the SAME model on the fused schedule — which forms the
    LayerNorm / add / GELU tensors first and quantises them with vq_act_quant_static, q|k|v through one shared pass —
    against the layer-by-layer schedule.  Image-sized token counts (4 frames x 256 tokens) so that the own attention kernels run."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from test_stdit_graph_cpu import Cfg
    from viditq_b200 import ops, ptq
    from viditq_b200.qdiff import QuantModel
    from viditq_b200.stdit import STDiT
    T, S = 4, 256
    model = STDiT(input_size=(T, 32, 32), depth=2)
    model.init_synthetic(seed=3)
    model.eval()
    sq = Cfg(enable=False, channel_wise_scale_type="momentum_act_max", momentum=0.95, alpha=0.625)
    wq = Cfg(n_bits=8, per_group="channel", channel_dim=0, scale_method="min_max", round_mode="nearest")
    aq = Cfg(n_bits=8, per_group=False, scale_method="min_max", round_mode="nearest_ste", running_stat=False, dynamic=False,
             sym=False, n_spatial_token=S, n_temporal_token=T, n_prompt=120, smooth_quant=sq)
    qnn = QuantModel(model, wq, aq)
    qnn.cuda()
    qnn.half()
    model.dtype = torch.float16
    g = torch.Generator().manual_seed(17)
    n = 4
    xs = torch.randn(n, 4, T, 32, 32, generator=g)
    ts = torch.tensor([800.0, 800.0, 200.0, 200.0])
    cs = torch.randn(n, 1, 120, 4096, generator=g).half()
    masks = torch.zeros(n, 120, dtype=torch.int64)
    masks[:, :77] = 1
    ptq.run_ptq(qnn, (xs, ts, cs, masks), n_samples=1, batch_size=1, fp_layer_list=FP_LAYERS, device=torch.device("cuda"))
    x = torch.randn(1, 4, T, 32, 32, generator=g).cuda()
    y = torch.randn(1, 1, 120, 4096, generator=g).cuda()
    mask = torch.zeros(1, 120, dtype=torch.int64)
    mask[0, :93] = 1
    mask = mask.cuda()
    t = torch.tensor([500.0], device="cuda")
    with torch.no_grad():
        n0 = ops.launch_count()
        ref = qnn(x, t, y, mask=mask).float().cpu()
        n1 = ops.launch_count()
        out = model.forward_fused(x, t, y, mask=mask).float().cpu()
        n2 = ops.launch_count()
        both = model.forward_fused(torch.cat([x, x]), torch.cat([t, t]), torch.cat([y, y]), mask=mask,
                                   independent=True).float().cpu()
    assert model._engine.static and ops.check_status() == 0 and torch.isfinite(out).all()
    assert all(v[1] for v in model._engine._static_same.values())       # q / k / v share their calibrated scales
    rel = ((out - ref).norm() / ref.norm()).item()
    print("static per-tensor scales: fused vs layer-by-layer schedule rel-L2 %.3e (%d vs %d own launches)"
          % (rel, n2 - n1, n1 - n0))
    assert rel <= 4e-3 and (n2 - n1) < (n1 - n0)
    assert torch.equal(both[0], out[0]) and torch.equal(both[1], out[0])      # static scales: stacking changes nothing
    # q / k / v with DIFFERENT calibrated scales (a hand-edited checkpoint): the shared pass must give way to three
    for blk in model.blocks:
        blk.attn.k.act_quantizer.delta.mul_(1.25)
        blk.attn_temp.v.act_quantizer.delta.mul_(0.8)
    with torch.no_grad():
        ref2 = qnn(x, t, y, mask=mask).float().cpu()
        out2 = model.forward_fused(x, t, y, mask=mask).float().cpu()
    assert not any(v[1] for v in model._engine._static_same.values())
    rel2 = ((out2 - ref2).norm() / ref2.norm()).item()
    print("static per-tensor scales, q / k / v scales differing: fused vs layer-by-layer rel-L2 %.3e" % rel2)
    assert rel2 <= 4e-3 and not torch.equal(out2, out)
