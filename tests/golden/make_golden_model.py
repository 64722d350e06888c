"""Generate tests/golden/stdit_small_golden.npz: the UNMODIFIED reference STDiT (2 blocks, hidden 1152, 4x16x16 latent)
wrapped in the reference QuantModel, W8A8 dynamic (w8a8_dynamic.yaml quantiser sections), run in fp16 on CPU.

Weights are the seeded synthetic init of viditq_b200.stdit.STDiT (state_dict-compatible, CPU generator => reproducible
on the GPU box without shipping 138 MB).  Stored: inputs, the reference's own quant ckpt (what ptq.py would save), its
W8A8 output and its fp16 un-quantised output.  Run here:  python tests/golden/make_golden_model.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install_opensora()
from opensora.models.stdit.stdit import STDiT as RefSTDiT  # noqa: E402
from qdiff.models.quant_model import QuantModel as RefQuantModel  # noqa: E402

from viditq_b200.stdit import STDiT  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stdit_small_golden.npz")
CFG = dict(input_size=(4, 16, 16), depth=2)
FP_LAYERS = ["x_embedder", "t_block", "t_embedder", "y_embedder", "final_layer"]   # remain_fp.txt


def main():
    torch.set_grad_enabled(False)
    mine = STDiT(**CFG)
    mine.init_synthetic(seed=0)
    ref = RefSTDiT(enable_flashattn=False, **CFG)
    ref.load_state_dict(mine.state_dict(), strict=True)
    ref.eval()
    T, S = ref.num_temporal, ref.num_spatial
    wq, aq = ref_shims.w8a8_dynamic_configs(n_temporal=T, n_spatial=S, n_prompt=120)
    wq["mixed_precision"] = [4, 6, 8]
    qnn = RefQuantModel(ref, wq, aq)
    qnn.set_module_name_for_quantizer(module=qnn.model)

    g = torch.Generator().manual_seed(1234)
    x = torch.randn(1, 4, 4, 16, 16, generator=g)
    y = torch.randn(1, 1, 120, 4096, generator=g).half().float()
    mask = torch.zeros(1, 120, dtype=torch.int64)
    mask[0, :77] = 1
    t = torch.tensor([500.0])

    # PTQ weight pass (ptq.py:266-294): fp32, weight quant on, act quant off, FP list kept FP
    qnn.set_quant_state(True, False)
    qnn.set_layer_quant(model=qnn, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")
    _ = qnn(x, t, y, mask=mask)
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    ckpt = qnn.get_quant_params_dict()
    rec = {"ckpt_names": np.array(sorted(ckpt.keys()))}
    for name, (bufs, params) in ckpt.items():
        assert len(params) == 0
        for bname, val in bufs.items():
            if val is not None:
                rec[f"ckpt/{name}/{bname}"] = val.detach().float().numpy()

    qnn.set_quant_state(False, False)
    out_fp32 = qnn(x, t, y, mask=mask)          # un-quantised fp32 graph (pins the host-side graph on CPU)
    rec["out_fp32"] = out_fp32.numpy()

    # inference state (quant_txt2video.py:195-207)
    qnn.set_quant_state(True, True)
    qnn.set_layer_quant(model=qnn, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")
    qnn.half()
    ref.dtype = torch.float16
    out_q = qnn(x, t, y, mask=mask)
    qnn.set_quant_state(False, False)
    out_fp = qnn(x, t, y, mask=mask)
    rec.update(x=x.numpy(), y=y.half().numpy(), mask=mask.numpy(), t=t.numpy(), out_w8a8=out_q.numpy(),
               out_fp16=out_fp.numpy(), T=np.int64(T), S=np.int64(S))
    np.savez_compressed(OUT, **rec)
    rel = (out_q - out_fp).norm() / out_fp.norm()
    print(f"wrote {OUT} ({os.path.getsize(OUT) / 1e6:.2f} MB); W8A8 vs fp16 rel-L2 = {rel:.4e}; |out| max {out_q.abs().max():.3f}")


if __name__ == "__main__":
    main()
