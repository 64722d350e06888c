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
