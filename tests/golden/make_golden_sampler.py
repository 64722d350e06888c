"""Generate tests/golden/sampler_golden.npz with the UNMODIFIED reference scheduler (t2v/opensora/schedulers/iddpm):
IDDPM(SpacedDiffusion).ddim_sample around forward_with_cfg (cfg_split=True) with a scripted stand-in denoiser.
The missing PTQD file ./t2v/rebuttal_files/k_for_each_timestep.pth is provided as zeros(20) in a temp cwd.
Run here:  python tests/golden/make_golden_sampler.py
"""
import os
import sys
import tempfile
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install_opensora()
import types  # noqa: E402
pk = types.ModuleType("opensora.schedulers")
pk.__path__ = [os.path.join(ref_shims.REFERENCE_ROOT, "t2v", "opensora", "schedulers")]
sys.modules["opensora.schedulers"] = pk
from opensora.schedulers.iddpm import IDDPM, forward_with_cfg  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sampler_golden.npz")


class Scripted:
    """Stand-in for QuantModel: cfg_split model whose two forwards return prescribed tensors."""
    cfg_split = True

    def __init__(self, out_c, out_u):
        self.outs = [out_c, out_u]
        self.calls = []

    def forward(self, x, t, y, **kw):
        self.calls.append((x.clone(), t.clone(), y.clone()))
        return self.outs[len(self.calls) - 1 if len(self.calls) <= 2 else (len(self.calls) - 1) % 2]


def main():
    rec = {}
    g = torch.Generator().manual_seed(7)
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "t2v", "rebuttal_files"))
        torch.save(torch.zeros(20), os.path.join(d, "t2v", "rebuttal_files", "k_for_each_timestep.pth"))
        cwd = os.getcwd()
        os.chdir(d)
        try:
            for steps, cfg in ((100, 4.0), (20, 7.0)):
                sch = IDDPM(num_sampling_steps=steps, cfg_scale=cfg)
                rec[f"s{steps}/timestep_map"] = np.array(sch.timestep_map)
                rec[f"s{steps}/alphas_cumprod"] = sch.alphas_cumprod
                rec[f"s{steps}/cfg_scale"] = np.float64(cfg)
                n = 1
                z = torch.randn(n, 4, 2, 8, 8, generator=g)
                x = torch.cat([z, z], 0)
                y = torch.randn(2 * n, 1, 6, 16, generator=g)
                for i in (steps - 1, steps // 2, 1, 0):
                    out_c = torch.randn(n, 8, 2, 8, 8, generator=g)
                    out_u = torch.randn(n, 8, 2, 8, 8, generator=g)
                    m = Scripted(out_c, out_u)
                    fwd = partial(forward_with_cfg, m, cfg_scale=cfg)
                    t = torch.tensor([i] * (2 * n))
                    res = sch.ddim_sample(fwd, x, t, clip_denoised=False, model_kwargs=dict(y=y))
                    rec[f"s{steps}/i{i}/x"] = z.numpy()
                    rec[f"s{steps}/i{i}/out_c"] = out_c.numpy()
                    rec[f"s{steps}/i{i}/out_u"] = out_u.numpy()
                    rec[f"s{steps}/i{i}/sample"] = res["sample"][:n].numpy()
                    rec[f"s{steps}/i{i}/model_t"] = m.calls[0][1].float().numpy()
        finally:
            os.chdir(cwd)
    np.savez_compressed(OUT, **rec)
    print("wrote", OUT, len(rec), "arrays")


if __name__ == "__main__":
    main()
