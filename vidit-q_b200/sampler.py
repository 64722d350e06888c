"""Caller side of the hot path: one ViDiT-Q denoising step = classifier-free-guidance forward(s) + DDIM update.

Restates, for the configuration the quant scripts use (IDDPM linear schedule, `num_sampling_steps` respacing, learned
sigma ignored by DDIM, eta = 0, cfg_split = True):
  * forward_with_cfg            t2v/opensora/schedulers/iddpm/__init__.py:135-184
  * SpacedDiffusion / ddim_sample  iddpm/respace.py:63-132, gaussian_diffusion.py:289-335, :514-552
The denoiser itself is whatever `model_forward(x, t, y, mask=...)` is (QuantModel.forward or STDiT.forward_fused).
"""
import numpy as np
import torch


def linear_beta_schedule(num_diffusion_timesteps=1000):
    scale = 1000 / num_diffusion_timesteps
    return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)


def space_timesteps(num_timesteps, section_count):
    """Evenly strided subset of timesteps (respace.py:8-60 for an integer section count)."""
    size = num_timesteps
    frac_stride = 1 if section_count <= 1 else (size - 1) / (section_count - 1)
    cur, steps = 0.0, []
    for _ in range(section_count):
        steps.append(round(cur))
        cur += frac_stride
    return sorted(set(steps))


class SpacedDDIM:
    def __init__(self, num_sampling_steps=100, diffusion_steps=1000, cfg_scale=4.0):
        betas = linear_beta_schedule(diffusion_steps)
        ac = np.cumprod(1.0 - betas)
        self.timestep_map = space_timesteps(diffusion_steps, num_sampling_steps)
        last, new_betas = 1.0, []
        for i in self.timestep_map:
            new_betas.append(1 - ac[i] / last)
            last = ac[i]
        betas = np.array(new_betas, dtype=np.float64)
        self.num_timesteps = len(betas)
        self.alphas_cumprod = np.cumprod(1.0 - betas)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.cfg_scale = cfg_scale

    def model_timestep(self, i):
        """Index in the spaced schedule -> timestep fed to the network (respace.py:121-132)."""
        return float(self.timestep_map[i])

    def coefficients(self, i, device):
        c = [self.sqrt_recip_alphas_cumprod[i], self.sqrt_recipm1_alphas_cumprod[i],
             np.sqrt(self.alphas_cumprod_prev[i]), np.sqrt(1 - self.alphas_cumprod_prev[i])]
        return torch.tensor(c, dtype=torch.float32, device=device)

    @staticmethod
    def cfg_combine(out_cond, out_uncond, cfg_scale, ptqd_k=0.0):
        """iddpm/__init__.py:166-184 on the first half of the (duplicated) batch: guidance on channels [:3] (sic),
        channels [3:] taken from the conditional branch; model_out / (1 + k) with k = 0 when no PTQD file exists."""
        c = out_cond / (1 + ptqd_k)
        u = out_uncond / (1 + ptqd_k)
        eps = u[:, :3] + cfg_scale * (c[:, :3] - u[:, :3])
        return torch.cat([eps, c[:, 3:]], dim=1)

    @staticmethod
    def ddim_update(x, model_out, coef):
        """gaussian_diffusion.py:289-335 + :540-552 with eta = 0: eps = first C channels (learned sigma unused)."""
        C = x.shape[1]
        eps = model_out[:, :C]
        pred_xstart = coef[0] * x - coef[1] * eps
        eps2 = (coef[0] * x - pred_xstart) / coef[1]
        return pred_xstart * coef[2] + coef[3] * eps2

    def step(self, model_forward, x, i, y_cond, y_uncond, mask):
        """One denoising step on latent x [n, C, T, H, W] (cfg_split: two forwards of batch n)."""
        t = torch.full((x.shape[0],), self.model_timestep(i), device=x.device)
        out_c = model_forward(x, t, y_cond, mask=mask)
        out_u = model_forward(x, t, y_uncond, mask=mask)
        out = self.cfg_combine(out_c, out_u, self.cfg_scale)
        return self.ddim_update(x, out, self.coefficients(i, x.device))
