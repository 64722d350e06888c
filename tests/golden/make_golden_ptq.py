"""Generate tests/golden/ptq_golden.npz: the reference's PTQ flow (t2v/scripts/ptq.py:213-362, smooth-quant branch) run with
the UNMODIFIED reference QuantModel / STDiT on a tiny model and a seeded synthetic calibration set — the act_scale
statistics and the per-timerange weight parameters ptq.py would save — for tests/test_ptq_cpu.py to hold
viditq_b200.ptq.run_ptq against.  Config: w4a8_timestep_aware_cb.yaml's quantiser sections (4-bit weights, mixed precision
[4, 6, 8], dynamic 8-bit activations, momentum smooth-quant with two timeranges), alphas made different per timerange.
Run here:  python tests/golden/make_golden_ptq.py
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

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ptq_golden.npz")
CFG = dict(input_size=(4, 16, 16), depth=2)
FP_LAYERS = ["x_embedder", "t_block", "t_embedder", "y_embedder", "final_layer"]   # remain_fp.txt
SMOOTH = dict(alpha=[0.11, 0.31], timerange=[[0, 500], [501, 1000]])
N_SAMPLES, BATCH_SIZE, STEPS = 2, 1, (900.0, 100.0)
SEED = 7


def calib_set():
    """(xs, ts, cond_embs, masks) in get_quant_calib_data's layout: per timestep 2 * n_samples entries, concatenated."""
    g = torch.Generator().manual_seed(99)
    n = 2 * N_SAMPLES * len(STEPS)
    xs = torch.randn(n, 4, 4, 16, 16, generator=g)
    ts = torch.tensor([t for t in STEPS for _ in range(2 * N_SAMPLES)])
    cs = torch.randn(n, 1, 120, 4096, generator=g).half().float()
    masks = torch.zeros(n, 120, dtype=torch.int64)
    for i in range(n):
        masks[i, :40 + 9 * i] = 1
    return xs, ts, cs, masks


def main():
    torch.set_grad_enabled(False)
    mine = STDiT(**CFG)
    mine.init_synthetic(seed=0)
    ref = RefSTDiT(enable_flashattn=False, **CFG)
    ref.load_state_dict(mine.state_dict(), strict=True)
    ref.eval()
    T, S = ref.num_temporal, ref.num_spatial
    wq, aq = ref_shims.w8a8_dynamic_configs(n_temporal=T, n_spatial=S, n_prompt=120, w_bits=4, smooth=SMOOTH)
    wq["mixed_precision"] = [4, 6, 8]
    qnn = RefQuantModel(ref, wq, aq)
    calib_xs, calib_ts, calib_cs, calib_masks = calib_set()
    calib_batch_size = BATCH_SIZE * 2
    tmp_kwargs = {"mask": calib_masks[:calib_batch_size][::2]}
    qnn.set_module_name_for_quantizer(module=qnn.model)

    # ---- ptq.py:219-262 verbatim in structure: smooth-quant statistics
    np.random.seed(SEED)
    qnn.set_smooth_quant(smooth_quant=False, smooth_quant_running_stat=True)
    qnn.set_quant_state(False, False)
    n_per = N_SAMPLES * 2
    ts = calib_ts.reshape([-1, n_per])
    n_steps = ts.shape[0]
    xs = calib_xs.reshape([n_steps, n_per] + list(calib_xs.shape[1:]))
    cs = calib_cs.reshape([n_steps, n_per] + list(calib_cs.shape[1:]))
    ms = calib_masks.reshape([n_steps, n_per] + list(calib_masks.shape[1:]))
    inds = np.arange(xs.shape[1])
    np.random.shuffle(inds)
    rounds = int(xs.size(1) / calib_batch_size)
    for i_ts in range(n_steps):
        for i in range(rounds):
            sel = inds[i * calib_batch_size:(i + 1) * calib_batch_size]
            _ = qnn(xs[i_ts, sel], ts[i_ts, sel], cs[i_ts, sel], mask=ms[i_ts, sel])
    qnn.set_smooth_quant(smooth_quant=True, smooth_quant_running_stat=False)
    qnn.set_layer_smooth_quant(model=qnn, module_name_list=FP_LAYERS, smooth_quant=False, smooth_quant_running_stat=False)

    # ---- ptq.py:264-294: weights, one calibration forward per timerange start
    qnn.set_quant_state(True, False)
    qnn.set_layer_quant(model=qnn, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")
    for range_start in [tr[0] for tr in SMOOTH["timerange"]]:
        _ = qnn(calib_xs[:calib_batch_size], calib_ts[:calib_batch_size].clone().fill_(range_start),
                calib_cs[:calib_batch_size], **tmp_kwargs)
    qnn.set_quant_init_done("weight")
    # ---- ptq.py:296-362: dynamic activations -> nothing to calibrate
    qnn.set_quant_state(True, True)
    qnn.set_layer_quant(model=qnn, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")
    qnn.set_quant_init_done("activation")
    ckpt = qnn.get_quant_params_dict()
    rec = {"ckpt_names": np.array(sorted(ckpt.keys())), "seed": np.int64(SEED)}
    n_buf = 0
    for name, (bufs, params) in ckpt.items():
        assert len(params) == 0
        for bname, val in bufs.items():
            if val is not None:
                rec[f"ckpt/{name}/{bname}"] = val.detach().float().numpy()
                n_buf += 1
    # ---- the same weight pass with the model in fp16 (ptq.py runs with dtype = "fp16" in the 16x512x512 config): the
    # reference then evaluates min / max / delta / zero point in HALF.  A fresh reference QuantModel takes the act_scale
    # statistics collected above (cast to half) and runs the two weight-calibration forwards in fp16.
    ref16 = RefSTDiT(enable_flashattn=False, **CFG)
    ref16.load_state_dict(mine.state_dict(), strict=True)
    ref16.eval()
    qnn16 = RefQuantModel(ref16, wq, aq)
    qnn16.set_module_name_for_quantizer(module=qnn16.model)
    qnn16.half()
    ref16.dtype = torch.float16
    src = dict(qnn.model.named_modules())
    for name, m in qnn16.model.named_modules():
        if hasattr(m, "act_quantizer") and hasattr(m.act_quantizer, "act_scale"):
            m.act_quantizer.act_scale = src[name].act_quantizer.act_scale.clone().half()
    qnn16.set_smooth_quant(smooth_quant=True, smooth_quant_running_stat=False)
    qnn16.set_layer_smooth_quant(model=qnn16, module_name_list=FP_LAYERS, smooth_quant=False, smooth_quant_running_stat=False)
    qnn16.set_quant_state(True, False)
    qnn16.set_layer_quant(model=qnn16, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                          act_quant=False, prefix="")
    for range_start in [tr[0] for tr in SMOOTH["timerange"]]:
        _ = qnn16(calib_xs[:calib_batch_size], calib_ts[:calib_batch_size].clone().fill_(range_start),
                  calib_cs[:calib_batch_size].half(), **tmp_kwargs)
    n16 = 0
    for name, (bufs, params) in qnn16.get_quant_params_dict().items():
        for bname, val in bufs.items():
            if val is not None and bname in ("delta_list", "zero_point_list"):
                assert val.dtype == torch.float16
                rec[f"ckpt16/{name}/{bname}"] = val.detach().numpy()
                n16 += 1
    # ---- static activation quantisers (w8a8_naive.yaml: `per_group: False, dynamic: False`), model in fp16 as ptq.py runs
    # it: weight pass, then ptq.py:311-327 — the calibration set walked in order with weights AND activations quantised,
    # every ActQuantizer re-initialising itself from each tensor it sees.  Two variants: last batch wins (running_stat
    # False, the shipped config) and the min / max EMA (running_stat True).
    for tag, running in (("static", False), ("static_ema", True)):
        refs = RefSTDiT(enable_flashattn=False, **CFG)
        refs.load_state_dict(mine.state_dict(), strict=True)
        refs.eval()
        wq_s, aq_s = ref_shims.w8a8_dynamic_configs(n_temporal=T, n_spatial=S, n_prompt=120)
        aq_s["dynamic"] = False
        aq_s["per_group"] = False
        aq_s["running_stat"] = running
        qs = RefQuantModel(refs, wq_s, aq_s)
        qs.set_module_name_for_quantizer(module=qs.model)
        qs.half()
        refs.dtype = torch.float16
        qs.set_quant_state(True, False)
        qs.set_layer_quant(model=qs, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                           act_quant=False, prefix="")
        _ = qs(calib_xs[:calib_batch_size], calib_ts[:calib_batch_size], calib_cs[:calib_batch_size].half(), **tmp_kwargs)
        qs.set_quant_init_done("weight")
        qs.set_quant_state(True, True)
        qs.set_layer_quant(model=qs, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                           act_quant=False, prefix="")
        for i in range(int(calib_xs.size(0) / calib_batch_size)):
            sel = slice(i * calib_batch_size, (i + 1) * calib_batch_size)
            _ = qs(calib_xs[sel], calib_ts[sel], calib_cs[sel].half(), mask=calib_masks[sel][::2])
        qs.set_quant_init_done("activation")
        n_s = 0
        for name, (bufs, params) in qs.get_quant_params_dict().items():
            for bname, val in bufs.items():
                if val is not None:
                    rec[f"{tag}/{name}/{bname}"] = val.detach().float().numpy()
                    n_s += 1
        print(f"{tag}: {n_s} buffers")
    np.savez_compressed(OUT, **rec)
    print(f"fp16 weight pass: {n16} buffers")
    print(f"wrote {OUT} ({os.path.getsize(OUT) / 1e6:.2f} MB): {len(ckpt)} quantisers, {n_buf} buffers")


if __name__ == "__main__":
    main()
