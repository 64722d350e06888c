"""Generate tests/golden/pixart_small_golden.npz: the UNMODIFIED reference PixArtMS (2 blocks, hidden 1152, 16x16 latent)
in the reference QuantModel(model_type="pixart"), W8A8 dynamic per-token, CFG batch-concat forward (batch 2: the
per-token statistics pool over the cond/uncond pair, quirk Q1), fp16 on CPU.  FP list of t2i/scripts/quant_txt2img.py:294
(final_layer stays quantised).  The stateful running-stat smooth-quant of blocks.27.mlp.fc2 (quirk Q17) is not enabled.
Run here:  python tests/golden/make_golden_pixart.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install_pixart()
from diffusion.model.nets.PixArtMS import PixArtMS as RefPixArt  # noqa: E402
from qdiff.models.quant_model import QuantModel as RefQuantModel  # noqa: E402

from viditq_b200.pixart import PixArtMS  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pixart_small_golden.npz")
CFG = dict(input_size=16, depth=2)
FP_LAYERS = ["x_embedder", "t_embedder", "t_block", "y_embedder", "csize_embedder", "ar_embedder"]


def main():
    torch.set_grad_enabled(False)
    mine = PixArtMS(**CFG)
    mine.init_synthetic(seed=0)
    ref = RefPixArt(**CFG)
    ref.load_state_dict(mine.state_dict(), strict=True)
    ref.eval()
    wq, aq = ref_shims.w8a8_dynamic_configs(n_temporal=1, n_spatial=64, n_prompt=120)
    wq["mixed_precision"] = [4, 6, 8]
    qnn = RefQuantModel(ref, wq, aq, model_type="pixart")
    qnn.set_module_name_for_quantizer(module=qnn.model)
    g = torch.Generator().manual_seed(4321)
    z = torch.randn(1, 4, 16, 16, generator=g)
    x = torch.cat([z, z], 0)
    y = torch.randn(2, 1, 120, 4096, generator=g).half().float()
    mask = torch.zeros(2, 120, dtype=torch.int64)
    mask[0, :93] = 1
    mask[1, :93] = 1
    t = torch.tensor([500.0, 500.0])

    def fp_list():
        qnn.set_layer_quant(model=qnn, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                            act_quant=False, prefix="")
    qnn.set_quant_state(True, False)
    fp_list()
    _ = qnn(x, t, y, mask=mask)
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    ckpt = qnn.get_quant_params_dict()
    rec = {"ckpt_names": np.array(sorted(ckpt.keys()))}
    for name, (bufs, params) in ckpt.items():
        for bname, val in bufs.items():
            if val is not None:
                rec[f"ckpt/{name}/{bname}"] = val.detach().float().numpy()
    qnn.set_quant_state(False, False)
    rec["out_fp32"] = qnn(x, t, y, mask=mask).numpy()
    qnn.set_quant_state(True, True)
    fp_list()
    qnn.half()
    # (PixArt.dtype is a property that follows the parameters)
    out_q = qnn(x, t, y, mask=mask)
    qnn.set_quant_state(False, False)
    out_fp = qnn(x, t, y, mask=mask)
    rec.update(x=x.numpy(), y=y.half().numpy(), mask=mask.numpy(), t=t.numpy(), out_w8a8=out_q.float().numpy(),
               out_fp16=out_fp.float().numpy())
    np.savez_compressed(OUT, **rec)
    rel = (out_q.float() - out_fp.float()).norm() / out_fp.float().norm()
    print(f"wrote {OUT} ({os.path.getsize(OUT) / 1e6:.2f} MB); W8A8 vs fp16 rel-L2 = {rel:.4e}")


if __name__ == "__main__":
    main()
