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
        # the reference extracts float64 tables to fp32 tensors and takes the square roots in fp32
        # (gaussian_diffusion.py:543-549 with sigma = 0)
        ab_prev = torch.tensor(self.alphas_cumprod_prev[i], dtype=torch.float64).float()
        c = torch.stack([torch.tensor(self.sqrt_recip_alphas_cumprod[i], dtype=torch.float64).float(),
                         torch.tensor(self.sqrt_recipm1_alphas_cumprod[i], dtype=torch.float64).float(),
                         torch.sqrt(ab_prev), torch.sqrt(1 - ab_prev)])
        return c.to(device)

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

    def step(self, model_forward, x, i, y_cond, y_uncond, mask, stacked_forward=None):
        """One denoising step on latent x [n, C, T, H, W] (cfg_split: two forwards of batch n).
        stacked_forward(x2, t2, y2, mask=) — e.g. functools.partial(STDiT.forward_fused, independent=True) — runs the
        cond and uncond calls as one stacked launch sequence; only used for n == 1, where "statistics pooled over the
        batch of each call" (quirk Q1) and "every row on its own" are the same thing."""
        t = torch.full((x.shape[0],), self.model_timestep(i), device=x.device)
        if stacked_forward is not None and x.shape[0] == 1:
            out = stacked_forward(torch.cat([x, x]), torch.cat([t, t]), torch.cat([y_cond, y_uncond]), mask=mask)
            out_c, out_u = out[:1], out[1:]
        else:
            out_c = model_forward(x, t, y_cond, mask=mask)
            out_u = model_forward(x, t, y_uncond, mask=mask)
        coef = self.coefficients(i, x.device)
        if x.is_cuda:   # fused sampler update (vq_cfg_ddim_step); the two static methods below are its restatement
            from . import ops
            return ops.cfg_ddim_step(out_c.float().contiguous(), out_u.float().contiguous(), x.float().contiguous(),
                                     coef, self.cfg_scale)
        out = self.cfg_combine(out_c, out_u, self.cfg_scale)
        return self.ddim_update(x, out, coef)


def get_key_for_value(dict_ranges, value):
    """gaussian_diffusion.py:24-29: the "hi-lo" range key containing step index `value` (non-range keys skipped)."""
    for key in dict_ranges:
        parts = str(key).split("-")
        if len(parts) != 2 or not all(p.isdigit() for p in parts):
            continue
        if int(parts[0]) >= value >= int(parts[1]):
            return key
    return None


class TimestepMixedPrecision:
    """Per-timestep bit-width switching of config 4 (quant_txt2video_mp.py:533-540 + gaussian_diffusion.py:739-759):
    whenever the DDIM step index enters a new range key, re-open the previous FP list, apply the range's FP list and
    load the per-layer weight / activation bit tables (QuantModel.load_bitwidth_config; delta stays, quirk Q7)."""

    def __init__(self, qnn):
        self.qnn = qnn
        self.key, self.fp_prev = None, None

    def before_step(self, i):
        qnn = self.qnn
        if not getattr(qnn, "timestep_wise_mp", False):
            return False
        key = get_key_for_value(qnn.time_mp_config_weight, i)
        if key is None:
            raise RuntimeError(f"this timestep {i} is not included by the config")
        if key == self.key:
            return False
        if self.fp_prev is not None:
            qnn.set_layer_quant(model=qnn, module_name_list=self.fp_prev, quant_level="per_layer", weight_quant=True,
                                act_quant=True, prefix="")
        fp_layers = qnn.time_mp_config_weight["fp_layers"][key]
        qnn.set_layer_quant(model=qnn, module_name_list=fp_layers, quant_level="per_layer", weight_quant=False,
                            act_quant=False, prefix="")
        qnn.load_bitwidth_config(model=qnn, bit_config=qnn.time_mp_config_weight[key], bit_type="weight")
        qnn.load_bitwidth_config(model=qnn, bit_config=qnn.time_mp_config_act[key], bit_type="act")
        self.key, self.fp_prev = key, fp_layers
        return True


def ddim_sample_loop(ddim: SpacedDDIM, model_forward, z, y_cond, y_uncond, mask, qnn=None, on_step=None,
                     stacked_forward=None):
    """iddpm IDDPM.sample(..., 'ddim') for cfg_split models: all steps from num_timesteps-1 down to 0."""
    mp = TimestepMixedPrecision(qnn) if qnn is not None else None
    for i in range(ddim.num_timesteps - 1, -1, -1):
        if mp is not None:
            mp.before_step(i)
        if qnn is not None:
            qnn.set_timestep_id_for_quantlayer(ddim.model_timestep(i))
        z = ddim.step(model_forward, z, i, y_cond, y_uncond, mask, stacked_forward=stacked_forward)
        if on_step is not None:
            on_step(i, z)
    return z


class GraphedSampler:
    """The complete DDIM sampling loop of one prompt (IDDPM.sample(..., 'ddim') with cfg_split) on the fused schedule, one
    CUDA-graph replay per step: the cond + uncond forwards run as ONE stacked launch sequence (un-pooled statistics == two
    batch-1 calls) followed by the fused CFG + DDIM update (vq_cfg_ddim_step), the latent is updated in place on the
    device, the timestep and the four DDIM coefficients are device tensors refreshed between replays.

    The kernels a step launches depend on which prepared weights it reads: one graph is captured per KEY =
    (smooth-quant timerange of the step's timestep, mixed-precision range of the step index) and reused for every step
    with that key — one graph for w8a8_dynamic.yaml, two for the two-timerange W4A8 config, one per range with
    per-timestep mixed precision (the per-(timerange, bit-config) graphs of SURVEY.md section 7 step 8)."""

    def __init__(self, qnn, model, ddim: SpacedDDIM, y_cond, y_uncond, mask, latent_shape):
        if not (y_cond.is_cuda and y_cond.shape[0] == 1 and latent_shape[0] == 1):
            raise ValueError("GraphedSampler runs one prompt (batch 1) per instance on a CUDA device")
        dev = y_cond.device
        self.qnn, self.model, self.ddim = qnn, model, ddim
        self.z = torch.zeros(latent_shape, device=dev)
        self.t = torch.zeros(1, device=dev)
        self.coef = torch.zeros(4, device=dev)
        self.y = torch.cat([y_cond, y_uncond]).contiguous()
        self.plan = model.mask_select_plan(mask.to(dev).repeat(2, 1) if mask.shape[0] == 1 else mask.to(dev))
        self.segments = model.kv_segments(self.plan[1], dev)
        self.mp = TimestepMixedPrecision(qnn)
        self.graphs = {}
        self.launches_per_step = 0

    def _step(self):
        from . import ops
        out = self.model.forward_fused(torch.cat([self.z, self.z]), self.t.expand(2), self.y, plan=self.plan,
                                       segments=self.segments, independent=True)
        self.z.copy_(ops.cfg_ddim_step(out[:1], out[1:], self.z, self.coef, self.ddim.cfg_scale))

    def _key(self, i):
        from .qdiff import find_interval
        t = self.ddim.model_timestep(i)
        tr = 0
        for _, layer in self.qnn.quant_layers():
            if getattr(layer, "smooth_quant", False) and hasattr(layer, "timerange"):
                tr = find_interval(layer.timerange, t)
                break
        rng = get_key_for_value(self.qnn.time_mp_config_weight, i) if getattr(self.qnn, "timestep_wise_mp", False) else None
        return tr, rng

    @torch.no_grad()
    def sample(self, z, on_step=None):
        """z: the initial noise [1, C, T, H, W] (any device) -> the final latent (a clone of the in-place buffer)."""
        self.z.copy_(z)
        for i in range(self.ddim.num_timesteps - 1, -1, -1):
            self.mp.before_step(i)
            t = self.ddim.model_timestep(i)
            self.qnn.set_timestep_id_for_quantlayer(t)
            self.t.fill_(t)
            self.coef.copy_(self.ddim.coefficients(i, "cpu"), non_blocking=True)
            key = self._key(i)
            g = self.graphs.get(key)
            if g is None:
                # a new key: run the step once eagerly on a copy of the state (prepares this key's weight codes, warms the
                # allocator), restore the latent, then capture
                keep = self.z.clone()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                from . import ops
                with torch.cuda.stream(side):
                    self._step()
                torch.cuda.current_stream().wait_stream(side)
                self.z.copy_(keep)
                g = torch.cuda.CUDAGraph()
                n0 = ops.launch_count()
                with torch.cuda.graph(g):
                    self._step()
                self.launches_per_step = ops.launch_count() - n0    # own kernels inside one captured step
                self.graphs[key] = g
                self.z.copy_(keep)      # the capture does not execute: the replay below is this step
            g.replay()
            if on_step is not None:
                on_step(i, self.z)
        return self.z.clone()
