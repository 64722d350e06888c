"""Generate tests/golden/stdit_deep_golden.npz: the UNMODIFIED reference at FULL DEPTH.

  * the reference STDiT-XL/2 graph (28 blocks, hidden 1152, 16 heads) on a small latent (4 x 16 x 16 -> T = 4, S = 64,
    256 tokens) wrapped in the reference QuantModel, W8A8 dynamic (w8a8_dynamic.yaml quantiser sections), fp16 on CPU:
    one forward (W8A8 and un-quantised fp16);
  * a 5-step DDIM sampling run through the reference scheduler — IDDPM(num_sampling_steps=5).ddim_sample_loop around
    forward_with_cfg with cfg_split = True (t2v/opensora/schedulers/iddpm/__init__.py:135-184,
    gaussian_diffusion.py:639-782) — storing the latent after every step, W8A8 and un-quantised fp16.

Weights are the seeded synthetic init of viditq_b200.stdit.STDiT (state_dict-compatible with the reference).  The
reference's PTQ weight pass produces the quant ckpt; the script ASSERTS that viditq_b200's
QuantModel.init_weight_quant_params() reproduces that ckpt exactly (delta_list, zero_point_list, delta, zero_point of
all 364 + 5 weight quantisers), so the 30 MB ckpt need not be committed: the tests rebuild it on the GPU box.
Run here (about 10 minutes, 8 CPU threads):  python tests/golden/make_golden_deep.py
"""
import os
import sys
import tempfile
import time
import types
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install_opensora()
pk = types.ModuleType("opensora.schedulers")
pk.__path__ = [os.path.join(ref_shims.REFERENCE_ROOT, "t2v", "opensora", "schedulers")]
sys.modules["opensora.schedulers"] = pk
from opensora.models.stdit.stdit import STDiT as RefSTDiT  # noqa: E402
from opensora.schedulers.iddpm import IDDPM, forward_with_cfg  # noqa: E402
from qdiff.models.quant_model import QuantModel as RefQuantModel  # noqa: E402

from viditq_b200.qdiff import QuantModel  # noqa: E402
from viditq_b200.stdit import STDiT  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stdit_deep_golden.npz")
CFG = dict(input_size=(4, 16, 16), depth=28)
FP_LAYERS = ["x_embedder", "t_block", "t_embedder", "y_embedder", "final_layer"]   # remain_fp.txt
N_STEPS, CFG_SCALE = 5, 4.0


def set_w8a8(qnn):
    qnn.set_quant_state(True, True)
    qnn.set_layer_quant(model=qnn, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")


def main():
    torch.set_grad_enabled(False)
    t_start = time.time()
    mine = STDiT(**CFG)
    mine.init_synthetic(seed=0)
    ref = RefSTDiT(enable_flashattn=False, **CFG)
    ref.load_state_dict(mine.state_dict(), strict=True)
    ref.eval()
    T, S = ref.num_temporal, ref.num_spatial
    wq, aq = ref_shims.w8a8_dynamic_configs(n_temporal=T, n_spatial=S, n_prompt=120)
    wq["mixed_precision"] = [4, 6, 8]
    qnn = RefQuantModel(ref, wq, aq)
    qnn.cfg_split = True
    qnn.set_module_name_for_quantizer(module=qnn.model)

    g = torch.Generator().manual_seed(4321)
    x = torch.randn(1, 4, 4, 16, 16, generator=g)
    y = torch.randn(1, 1, 120, 4096, generator=g).half().float()
    y_null = torch.randn(1, 1, 120, 4096, generator=g).half().float()
    mask = torch.zeros(1, 120, dtype=torch.int64)
    mask[0, :83] = 1
    t = torch.tensor([500.0])

    # PTQ weight pass (ptq.py:266-294): fp32, weight quant on, act quant off, FP list kept FP
    qnn.set_quant_state(True, False)
    qnn.set_layer_quant(model=qnn, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")
    _ = qnn(x, t, y, mask=mask)
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    ckpt = qnn.get_quant_params_dict()
    print(f"[{time.time() - t_start:.0f}s] reference weight pass done: {len(ckpt)} quantisers")

    # viditq_b200's own min-max weight init must reproduce the reference ckpt bit for bit (then the tests need no ckpt)
    wq2, aq2 = ref_shims.w8a8_dynamic_configs(n_temporal=T, n_spatial=S, n_prompt=120)
    wq2["mixed_precision"] = [4, 6, 8]
    q2 = QuantModel(mine, wq2, aq2)
    q2.set_module_name_for_quantizer(module=q2.model)
    q2.init_weight_quant_params()
    n_checked = 0
    for name, layer in q2.quant_layers():
        key = name + ".weight_quantizer"
        if not name.startswith("blocks."):
            continue        # FP-list layers: the reference's weight pass never initialises them (weight_quant False)
        bufs = ckpt[key][0]
        w2 = layer.weight_quantizer
        for b in ("delta_list", "zero_point_list", "delta", "zero_point"):
            a_, b_ = getattr(w2, b).float(), bufs[b].float()
            assert a_.shape == b_.shape and torch.equal(a_, b_), (key, b, (a_ - b_).abs().max())
        n_checked += 1
    assert n_checked == 13 * 28, n_checked
    print(f"[{time.time() - t_start:.0f}s] init_weight_quant_params == reference ckpt on {n_checked} layers (bit-exact)")
    del q2

    # inference state (quant_txt2video.py:195-207)
    set_w8a8(qnn)
    qnn.half()
    ref.dtype = torch.float16
    out_q = qnn(x, t, y, mask=mask)
    qnn.set_quant_state(False, False)
    out_fp = qnn(x, t, y, mask=mask)
    rel = ((out_q - out_fp).norm() / out_fp.norm()).item()
    print(f"[{time.time() - t_start:.0f}s] forward: W8A8 vs fp16 rel-L2 = {rel:.4e}; |out| max {out_q.abs().max():.3f}")
    rec = dict(x=x.numpy(), y=y.half().numpy(), y_null=y_null.half().numpy(), mask=mask.numpy(), t=t.numpy(),
               out_w8a8=out_q.numpy(), out_fp16=out_fp.numpy(), T=np.int64(T), S=np.int64(S),
               n_steps=np.int64(N_STEPS), cfg_scale=np.float64(CFG_SCALE), quant_err_forward=np.float64(rel))

    # 5-step DDIM through the reference scheduler, cfg_split (two batch-1 forwards per step)
    sch = IDDPM(num_sampling_steps=N_STEPS, cfg_scale=CFG_SCALE)
    rec["timestep_map"] = np.array(sch.timestep_map)
    z0 = torch.randn(1, 4, 4, 16, 16, generator=g)
    rec["z0"] = z0.numpy()
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "t2v", "rebuttal_files"))
        torch.save(torch.zeros(20), os.path.join(d, "t2v", "rebuttal_files", "k_for_each_timestep.pth"))
        cwd = os.getcwd()
        os.chdir(d)
        try:
            for tag, (wqs, aqs) in (("w8a8", (True, True)), ("fp16", (False, False))):
                if wqs:
                    set_w8a8(qnn)
                else:
                    qnn.set_quant_state(False, False)
                z = torch.cat([z0, z0], 0)
                kwargs = dict(y=torch.cat([y, y_null], 0), mask=mask)
                fwd = partial(forward_with_cfg, qnn, cfg_scale=CFG_SCALE)
                traj = []
                for out in sch.ddim_sample_loop_progressive(fwd, z.shape, noise=z, clip_denoised=False,
                                                            model_kwargs=kwargs, device="cpu", progress=False):
                    traj.append(out["sample"][:1].float().numpy())
                    print(f"[{time.time() - t_start:.0f}s] {tag} step {len(traj)}/{N_STEPS} |z| max "
                          f"{np.abs(traj[-1]).max():.3f}")
                rec[f"traj_{tag}"] = np.stack(traj)
        finally:
            os.chdir(cwd)
    d = rec["traj_w8a8"][-1].astype(np.float64) - rec["traj_fp16"][-1].astype(np.float64)
    rec["quant_err_sampling"] = np.float64(np.linalg.norm(d) / np.linalg.norm(rec["traj_fp16"][-1]))
    np.savez_compressed(OUT, **rec)
    print(f"wrote {OUT} ({os.path.getsize(OUT) / 1e6:.2f} MB); quantisation error after {N_STEPS} steps "
          f"rel-L2 = {rec['quant_err_sampling']:.4e}")


if __name__ == "__main__":
    main()
