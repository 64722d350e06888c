"""PTQ producer: what the reference's t2v/scripts/ptq.py:213-362 does between "calibration data
loaded" and "ckpt.pth saved": the dynamic-activation configurations (w8a8_dynamic.yaml, the headline; the smooth-quant /
timestep-aware variants w4a8_timestep_aware_cb.yaml, w8a8_smooth_quant.yaml with `dynamic: True`) and the static ones
(w8a8_naive.yaml).  The result is the reference's checkpoint format: `QuantModel.get_quant_params_dict()` — loadable by the
reference's own `load_quant_params` (qdiff/utils.py:66) and by `viditq_b200.qdiff.load_quant_params`.

Steps, in the reference's order:
  1. smooth-quant statistics (ptq.py:219-262): with quantisation off and `smooth_quant_running_stat` on, the calibration
     set is walked timestep by timestep in shuffled mini-batches; every QuantLayer keeps the EMA of its per-channel |x|
     maxima per timerange in `act_quantizer.act_scale` (quant_layer.py:146-153).  Here the forward runs on the device the
     model lives on; the per-channel maxima come from the `vq_col_absmax` kernel.
  2. weight parameters (ptq.py:264-294): min-max per output channel of W (or of W * channel_wise_scale per timerange
     under smooth-quant), every bit-width of `mixed_precision` — `QuantModel.init_weight_quant_params`, evaluated on the
     device without the reference's calibration forwards (the parameters depend on the weights and act_scale only;
     bit-exact against the reference's ckpt on 364 layers, tests/test_gpu_deep.py, and under smooth-quant,
     tests/test_ptq_cpu.py).
  3. activation parameters (ptq.py:296-362): dynamic quantisers have none ("Adopting dynamic quant params, skip
     calculating fixed quant params", ptq.py:308-309).  STATIC quantisers (w8a8_naive.yaml: per tensor; or per token) are
     calibrated by walking the calibration set with weights AND activations quantised (ptq.py:311-327): every call
     re-derives (delta, zero_point) from the live tensor (base_quantizer.py:146-228; min / max EMA with `running_stat`)
     and the forward continues on the quantised activations — here through the integer kernels.  Timestep-wise static
     parameters (ptq.py:328-356) raise.
Not part of the denoising hot path: plain torch on the model's device plus one own kernel; nothing here is timed by
bench.py.
"""
from collections import OrderedDict

import numpy as np
import torch

from .qdiff import _cfg_get, _is_dynamic


def _smooth_cfg(qnn):
    for _, layer in qnn.quant_layers():
        return _cfg_get(_cfg_get(layer.act_quant_params, "smooth_quant"), "enable", False)
    return False


@torch.no_grad()
def collect_smooth_quant_statistics(qnn, calib, n_samples, batch_size, fp_layer_list=(), device=None):
    """ptq.py:219-262.  calib = (xs, ts, cond_embs, masks) as qdiff.utils.get_quant_calib_data returns them: the selected
    timesteps concatenated along the batch, 2 * n_samples entries (cond + uncond) per timestep.  Uses numpy's global RNG
    for the mini-batch order exactly like the reference (seed it for reproducibility)."""
    xs, ts, cs, masks = calib
    calib_batch_size = batch_size * 2                       # "used to support the CFG", ptq.py:183
    qnn.set_smooth_quant(smooth_quant=False, smooth_quant_running_stat=True)
    qnn.set_quant_state(False, False)
    n_per_step = n_samples * 2
    ts = ts.reshape([-1, n_per_step])
    n_steps = ts.shape[0]
    xs = xs.reshape([n_steps, n_per_step] + list(xs.shape[1:]))
    cs = cs.reshape([n_steps, n_per_step] + list(cs.shape[1:]))
    masks = masks.reshape([n_steps, n_per_step] + list(masks.shape[1:]))
    inds = np.arange(xs.shape[1])
    np.random.shuffle(inds)
    rounds = int(xs.size(1) / calib_batch_size)
    to = (lambda v: v.to(device)) if device is not None else (lambda v: v)
    for i_ts in range(n_steps):
        assert torch.all(ts[i_ts, :] == ts[i_ts, 0])         # one timestep per group
        for i in range(rounds):
            sel = inds[i * calib_batch_size:(i + 1) * calib_batch_size]
            qnn(to(xs[i_ts, sel]), to(ts[i_ts, sel]), to(cs[i_ts, sel]), mask=to(masks[i_ts, sel]))
    qnn.set_smooth_quant(smooth_quant=True, smooth_quant_running_stat=False)
    qnn.set_layer_smooth_quant(model=qnn, module_name_list=list(fp_layer_list), smooth_quant=False,
                               smooth_quant_running_stat=False)


@torch.no_grad()
def calibrate_static_activations(qnn, layers, calib, batch_size, device=None):
    """ptq.py:311-327 (`timestep_wise: False`): with the weights quantised on their final grid, the calibration set is walked
    in order in mini-batches of 2 * batch_size; every static ActQuantizer re-derives its parameters from each tensor it sees
    (QuantLayer.calibrate_static_act: last batch wins, or the min / max EMA with `running_stat`) and the forward continues
    on the activations quantised with them — here through the integer kernels (vq_act_quant_static + vq_gemm_w8a8), which
    equal the reference's simulated quantisation to <= 1e-3 per layer."""
    if getattr(qnn, "timestep_wise", False):
        raise NotImplementedError("timestep-wise static activation parameters (ptq.py:328-356) are not restated")
    xs, ts, cs, masks = calib
    calib_batch_size = batch_size * 2
    to = (lambda v: v.to(device)) if device is not None else (lambda v: v)
    for layer in layers:
        layer.calibrating = True
        aq = layer.act_quantizer
        aq.init_done = False
        aq.x_min = aq.x_max = None                    # a fresh calibration: no EMA state, no parameters of an earlier run
        aq.delta_list = aq.zero_point_list = aq.delta = aq.zero_point = None
    try:
        for i in range(int(xs.size(0) / calib_batch_size)):
            sel = slice(i * calib_batch_size, (i + 1) * calib_batch_size)
            qnn(to(xs[sel]), to(ts[sel]), to(cs[sel]), mask=to(masks[sel][::2]))
    finally:
        for layer in layers:
            layer.calibrating = False


@torch.no_grad()
def run_ptq(qnn, calib, n_samples, batch_size, fp_layer_list=(), device=None):
    """The training-free PTQ of ptq.py:213-362 on `qnn` (a viditq_b200.qdiff.QuantModel); returns the checkpoint dict
    (`torch.save` it as ckpt.pth).  Leaves the model in the inference state of quant_txt2video.py:195-207: weights and
    activations quantised, the `fp_layer_list` layers in floating point."""
    qnn.set_module_name_for_quantizer(module=qnn.model)
    fp_layer_list = list(fp_layer_list)
    if _smooth_cfg(qnn):
        collect_smooth_quant_statistics(qnn, calib, n_samples, batch_size, fp_layer_list, device)
    # ---- weights (ptq.py:264-294, the part_fp branch)
    qnn.set_quant_state(True, False)
    qnn.set_layer_quant(model=qnn, module_name_list=fp_layer_list, quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")
    qnn.init_weight_quant_params(only_enabled=True, dtype=None)     # the model's own dtype, as the reference computes
    qnn.set_quant_init_done("weight")
    # ---- activations (ptq.py:296-362)
    qnn.set_quant_state(True, True)
    qnn.set_layer_quant(model=qnn, module_name_list=fp_layer_list, quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")
    static = [layer for _, layer in qnn.quant_layers() if layer.act_quant and not _is_dynamic(layer.act_quantizer)]
    if static:
        calibrate_static_activations(qnn, static, calib, batch_size, device)
    qnn.set_quant_init_done("activation")
    return qnn.get_quant_params_dict()


def save_ckpt(quant_params_dict, path):
    """ptq.py:400-404: the dict of (buffers, parameters) per quantiser, as torch.save writes it."""
    torch.save(OrderedDict(quant_params_dict), path)
