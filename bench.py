#!/usr/bin/env python
"""bench.py — STDiT 16x512x512 W8A8 denoise-steps/sec on N B200s (BASELINE.json metric), one JSON line.

A "step" is one full ViDiT-Q denoising step of one sample: cfg_split => two STDiT-XL/2 forwards (cond / uncond, 28
blocks, 16384 tokens) on the fused sm_100a kernels + CFG combine + DDIM update.  Synthetic latents / text embeddings,
seeded random-init weights, min-max weight scales.  One process per GPU; ranks hold independent samples (weak scaling,
no data-path collective — SURVEY.md §8e); value = steps of all ranks / max-over-ranks device time.

  python bench.py [--gpus N --steps K --warmup W]           the metric (BASELINE config 3 / 5 as a step rate)
  python bench.py --schedule hook[-graph]                      same step through the reference's hook API (QuantModel.forward:
                                                             one QuantLayer call per linear, torch SDPA attention)
  python bench.py --workload linear | pixart512 | w4a8mp     BASELINE configs 1, 2, 4
  python bench.py --impl reference ...                       the reference's simulated-quant path on the host cores (port)
"""
import argparse
import collections
import ctypes
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "stdit_16x512x512_w8a8_denoise_steps_per_sec"
UNIT = "steps/s"
T_FRAMES, S_TOKENS, HIDDEN, DEPTH, HEADS, PROMPT_LEN = 16, 1024, 1152, 28, 16, 120
FP_LAYERS = ["x_embedder", "t_block", "t_embedder", "y_embedder", "final_layer"]   # remain_fp.txt
FP_LAYERS_PIXART = ["x_embedder", "t_embedder", "t_block", "y_embedder", "csize_embedder", "ar_embedder"]
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")   # written by tools/ncu_traffic.py from an ncu capture


def linear_ops_per_forward(n_tok=T_FRAMES * S_TOKENS, L=PROMPT_LEN):
    C = HIDDEN
    per_block = 10 * n_tok * C * C + 2 * n_tok * C * 4 * C + L * C * 2 * C
    return 2 * per_block * DEPTH


def pixart_linear_ops(depth=28, M=2048, L=2 * 109):
    """PixArt-alpha 512 solver step (CFG batch 2): qkv, proj, q_linear, cross proj, fc1, fc2 on M = 2048 rows, kv_linear on the
    prompt rows, per block; + the quantised final_layer.linear (1152 -> 32).  SURVEY.md §8: 2.17 TOP at 28 blocks."""
    C = HIDDEN
    per_block = M * C * (3 * C + C + C + C + 4 * C + 4 * C) + L * C * 2 * C
    return 2.0 * (depth * per_block + M * C * 32)


class Cfg(dict):
    __getattr__ = dict.get


def quant_cfgs(w_bits=8, smooth=None, n_spatial=S_TOKENS, n_temporal=T_FRAMES, static=False):
    sq = Cfg(enable=False, channel_wise_scale_type="momentum_act_max", momentum=0.95, alpha=0.625)
    if smooth is not None:
        sq = Cfg(enable=True, channel_wise_scale_type="momentum_act_max", momentum=0.95, **smooth)
    wq = Cfg(n_bits=w_bits, per_group="channel", channel_dim=0, scale_method="min_max", round_mode="nearest",
             mixed_precision=[4, 6, 8])
    aq = Cfg(n_bits=8, per_group="token", scale_method="min_max", round_mode="nearest_ste", running_stat=False,
             dynamic=True, sym=False, n_spatial_token=n_spatial, n_temporal_token=n_temporal, n_prompt=PROMPT_LEN,
             smooth_quant=sq)
    if static:      # w8a8_naive.yaml: one calibrated (delta, zero point) per activation tensor
        aq["per_group"], aq["dynamic"] = False, False
    return wq, aq


# --------------------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md)
# --------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.samples, self._stop, self._th, self._proc = gpu_index, [], threading.Event(), None, None
        self.t0 = self.t1 = None   # host-time window of the timed region (device work is bracketed by synchronize)

    def _run(self):
        # one streaming nvidia-smi (-lms 50): a fresh process per sample costs > 100 ms and would see only a couple of
        # samples of a sub-second timed region
        try:
            self._proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.FIELDS}",
                                           "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                          stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        for line in self._proc.stdout:
            parts = [p.strip() for p in line.strip().split(",")]
            if len(parts) >= 6 and self.t0 is not None and self.t1 is None:   # inside the timed region only
                self.samples.append(parts)
            if self._stop.is_set():
                break

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        proc = getattr(self, "_proc", None)
        if proc is not None:
            proc.terminate()       # the exact child we started
        self._th.join(timeout=10)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        mx = max((int(s[1]) for s in self.samples if s[1].isdigit()), default=None)
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's simulated-quant path (oracle/cpu_reference_arm.py — the Python reference cannot travel)
# --------------------------------------------------------------------------------------------------------------------
def cpu_baseline(n_blocks=2):
    from oracle import cpu_reference_arm as R
    R.time_sample(1)                       # warm-up: page in the weights, size the thread pool
    s = R.time_sample(n_blocks)
    return {"value": R.steps_per_sec(s), "unit": UNIT, "cores": s["cores"], "kind": "port", "sample": R.describe(s)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_reference_arm as R
    for _ in range(min(args.warmup, 2)):
        R.time_sample(1)
    samples = [R.time_sample(1) for _ in range(max(1, min(args.steps, 8)))]
    best = min(samples, key=lambda s: s["seconds"])
    mean_block = sum(s["block_seconds"] for s in samples) / len(samples)
    mean_rest = sum(s["rest_seconds"] for s in samples) / len(samples)
    val = 1.0 / (2.0 * (mean_rest + DEPTH * mean_block))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / val, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "STDiT-XL/2 16x512x512 W8A8 (w8a8_dynamic), cfg_split, DDIM step",
                       "note": "reference = the simulated-quant CPU path (fake-quant in fp32 + F.linear), torch port pinned "
                               "bit-exact to the reference classes; each timed step is a bounded sample (embed + 1 full-size "
                               "block + final layer of one forward, %d samples), a denoise step = 2 forwards x 28 blocks is "
                               "extrapolated linearly" % len(samples)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": best["cores"], "kind": "port", "sample": R.describe(best)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------------
# measurement helpers
# --------------------------------------------------------------------------------------------------------------------
def measure_int8_peak(dev):
    """Dense INT8 tensor-pipe rate of THIS GPU, measured with this repo's own mainloop: vq_gemm_w8a8 of the measurement build
    (libviditq_b200_dbg.so, -DVQ_DEBUG_EPI) in its mainloop-only mode (TMA -> tcgen05.mma.kind::i8 -> accumulators discarded,
    no epilogue) on 16384 x 4608 x 4608.  burst = best of 10 single launches, sustained = back to back for 2 s.
    Returns None when the measurement build is absent."""
    import torch
    path = os.path.join(ROOT, "vidit-q_b200", "libviditq_b200_dbg.so")
    if not os.path.exists(path):
        return None
    L = ctypes.CDLL(path)
    vp, i32 = ctypes.c_void_p, ctypes.c_int
    L.vq_gemm_w8a8.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, i32, i32, i32, vp, i32, vp, i32, vp, i32, vp]
    L.vq_gemm_w8a8.restype = i32
    M, N, K = 16384, 4608, 4608
    a = torch.randint(0, 255, (M, K), dtype=torch.uint8, device=dev)
    w = torch.randint(0, 255, (N, K), dtype=torch.uint8, device=dev)
    dl = torch.ones(M, dtype=torch.float16, device=dev)
    zp = torch.zeros(M, dtype=torch.float16, device=dev)
    rs = torch.zeros(M, dtype=torch.int32, device=dev)
    col = torch.zeros(N, 4, dtype=torch.int32, device=dev)
    out = torch.empty(M, N, dtype=torch.float16, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def launch():
        rc = L.vq_gemm_w8a8(a.data_ptr(), dl.data_ptr(), zp.data_ptr(), rs.data_ptr(), M, w.data_ptr(), col.data_ptr(), M, N,
                            K, 3, None, N, None, 0, out.data_ptr(), N, st)
        if rc != 0:
            raise RuntimeError(f"mainloop-only GEMM failed: {rc}")
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    ops = 2.0 * M * N * K
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    n, t_start = 0, time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t_start < 2.0:
        for _ in range(50):
            launch()
        n += 50
        torch.cuda.current_stream().synchronize()
    e1.record()
    torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / n
    return {"int8_tops_burst": ops / (best * 1e-3) / 1e12, "int8_tops_sustained": ops / (sus * 1e-3) / 1e12,
            "how": "vq_gemm_w8a8 mainloop-only (measurement build), 16384x4608x4608 u8: best of 10 / back to back for 2 s"}


KERNEL_CLASSES = [
    # (class, regex on the kernel name, bound)
    ("gemm", r"vq_gemm_w8a8_kernel|vq_linear_fused_kernel", "tensor"),
    ("quant", r"vq_act_quant|vq_col_absmax|ia_stats_kernel|ia_quant_kernel", "hbm"),
    ("attn_tc", r"vq_attn_spatial_kernel|vq_attn_i8_kernel", "tensor"),
    ("attn_small", r"vq_attn_temporal|vq_attn_cross_kernel", "hbm"),
    ("embed_sampler", r"vq_patch_embed_kernel|vq_cfg_ddim_kernel", "hbm"),
]


def profile_kernels(run):
    """CUPTI (torch.profiler) timeline of ONE execution of `run` — for the CUDA-graph arm the replayed graph itself, so the
    per-kernel durations are the ones inside the timed step, not those of an instrumented eager pass.
    Returns {class: (launches, total_ms)}, total busy ms, total span ms."""
    import torch
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    agg = collections.defaultdict(lambda: [0, 0.0])
    busy = 0.0
    for e in evs:
        dur = (e.time_range.end - e.time_range.start) / 1e3
        busy += dur
        for cls, pat, _ in KERNEL_CLASSES:
            if re.search(pat, e.name):
                agg[cls][0] += 1
                agg[cls][1] += dur
                break
        else:
            agg["other"][0] += 1
            agg["other"][1] += dur
    span = (max(e.time_range.end for e in evs) - min(e.time_range.start for e in evs)) / 1e3 if evs else 0.0
    return {k: (v[0], v[1]) for k, v in agg.items()}, busy, span


class WorkMeter:
    """Algorithmic work of one step, counted from the arguments of the kernel wrappers while an eager step runs (no literals):
    class -> [launches, ops, bytes].  bytes = the tensors a launch must read and write once."""

    def __init__(self, ops):
        self.ops, self.saved, self.work = ops, {}, collections.defaultdict(lambda: [0, 0.0, 0.0])

    def add(self, cls, launches, ops_, bytes_):
        w = self.work[cls]
        w[0] += launches
        w[1] += ops_
        w[2] += bytes_

    def __enter__(self):
        o, me = self.ops, self

        def wrap(name, fn):
            orig = getattr(o, name)
            me.saved[name] = orig

            def inner(*a, **k):
                fn(*a, **k)
                return orig(*a, **k)
            setattr(o, name, inner)

        def gemm(a, w, epi=0, **k):
            M = a.G * a.rows
            me.add("gemm", 1, 2.0 * M * w.N * w.K, M * w.K + w.N * w.K + 2 * M * w.N * (2 if epi == 2 else 1))

        def linear(x, w, epi=0, **k):
            G, rows, K = x.shape
            M = G * rows
            n = o.linear_launch_count(G, rows, K)
            if n == 1:
                me.add("gemm", 1, 2.0 * M * w.N * K, 2 * M * K + w.N * K + 2 * M * w.N * (2 if epi == 2 else 1))
            else:
                me.add("quant", 1, 0, 3 * M * K + 8 * M)
                me.add("gemm", 1, 2.0 * M * w.N * K, M * K + w.N * K + 2 * M * w.N * (2 if epi == 2 else 1))

        def quant(x, *a, **k):
            M, K = x.numel() // x.shape[-1], x.shape[-1]
            me.add("quant", 1, 0, 3 * M * K + 8 * M)
        wrap("gemm_w8a8", gemm)
        wrap("linear_w8a8", linear)
        for name in ("act_quant", "add_act_quant", "act_quant_static", "act_quant_heads", "ln_modulate_act_quant"):
            wrap(name, quant)
        wrap("col_absmax", lambda x, **k: me.add("quant", 1, 0, 2 * x.numel()))
        wrap("attn_spatial", lambda qkv, n_seq, S, H, D, scale, **k: me.add("attn_tc", 1, 4.0 * n_seq * H * S * S * D, 8 * n_seq * S * H * D))
        # opt-in INT8 attention: the same 4 S^2 D operations per head (now integer), plus the two operand passes (statistics:
        # read k|v; codes: read q|k|v, write the 80-byte Q8/K8 rows and V8^T)
        wrap("attn_spatial_i8", lambda qkv, n_seq, S, H, D, scale, **k: (
            me.add("attn_tc", 1, 4.0 * n_seq * H * S * S * D, n_seq * S * H * (2 * 80 + D) + 2 * n_seq * S * H * D),
            me.add("quant", 2, 0, n_seq * S * H * D * (4 + 6) + n_seq * S * H * (2 * 80 + D))))
        wrap("attn_cross", lambda q, kv, ks, kl, B, N, H, D, max_len, scale, **k: me.add(
            "attn_tc" if N % 256 == 0 else "attn_small", 1, 4.0 * N * kv.shape[0] * H * D, 4 * B * N * H * D + 2 * kv.numel()))
        wrap("attn_temporal", lambda qkv, B, T, S, H, D, scale, **k: me.add("attn_small", 1, 4.0 * B * S * H * T * T * D, 8 * B * T * S * H * D))
        wrap("attn_temporal_quant", lambda qkv, B, T, S, H, D, scale, **k: me.add("attn_small", 1, 4.0 * B * S * H * T * T * D,
                                                                                    7 * B * T * S * H * D + 8 * B * T * S))
        wrap("patch_embed", lambda latent, weight, *a, **k: me.add("embed_sampler", 1, 0, 4 * latent.numel() + 2 * (latent.numel() // (latent.shape[1] * 4)) * weight.shape[0]))
        wrap("cfg_ddim_step", lambda oc, ou, x, *a, **k: me.add("embed_sampler", 1, 0, 8 * oc.numel() + 8 * x.numel()))
        return self

    def __exit__(self, *a):
        for name, orig in self.saved.items():
            setattr(self.ops, name, orig)


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def load_traffic():
    try:
        return json.load(open(TRAFFIC_FILE))
    except Exception:
        return {}


def rooflines(work, prof, peaks, int8_peak, busy_ms, ms_per_step_timed):
    """One roofline entry per kernel class: achieved = algorithmic ops (tensor-bound) or bytes (HBM-bound) of the class over
    one step / the class's kernel time inside the TIMED step.  That time = the class's SHARE of the CUPTI-profiled replay x
    the step time of the timed region: a single profiled replay runs on a cool GPU at higher clocks than K back-to-back
    steps under the 1000 W cap, so its absolute durations would overstate what the timed region achieved."""
    steps_profiled = 1
    hbm = peaks.get("hbm_gbs", 6500.0)
    bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
    traffic = load_traffic()
    out = []
    for cls, _, bound in KERNEL_CLASSES:
        if cls not in prof or cls not in work:
            continue
        launches, ms_raw = prof[cls]
        ms = ms_raw / busy_ms * ms_per_step_timed if busy_ms > 0 else ms_raw
        n, ops_, bytes_ = work[cls]
        if ms <= 0:
            continue
        tr = traffic.get(cls)
        entry = {"kernel_class": cls, "bound": bound, "launches_per_step": launches // steps_profiled,
                 "ms_per_step": ms, "ms_per_step_single_profiled_replay": ms_raw, "share_of_step": ms_raw / busy_ms,
                 "algorithmic_bytes_per_launch": bytes_ / max(1, n),
                 "traffic": None if tr is None else tr.get("dram_bytes_per_launch"),
                 "traffic_source": None if tr is None else tr.get("source")}
        if bound == "tensor":
            if cls == "gemm":
                if int8_peak:
                    peak, src = int8_peak["int8_tops_sustained"], ("measured: " + int8_peak["how"] +
                                                                     f" (burst {int8_peak['int8_tops_burst']:.0f} TOP/s)")
                else:
                    peak, src = 2.0 * bf16, "proxy: 2 x bf16_tflops_sustained of MEASURED_PEAKS.json (measurement build absent)"
                unit = "TOP/s"
            else:
                peak, src, unit = bf16, "bf16_tflops_sustained of MEASURED_PEAKS.json (fp16 attention, sustained: timed inside a long step)", "TFLOP/s"
            ach = ops_ / (ms * 1e-3) / 1e12
        else:
            peak, src, unit = hbm, "hbm_gbs of MEASURED_PEAKS.json" if peaks else "fallback 6500 GB/s", "GB/s"
            ach = bytes_ / (ms * 1e-3) / 1e9
        entry.update({"achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "peak_source": src})
        out.append(entry)
    return out


# --------------------------------------------------------------------------------------------------------------------
# models
# --------------------------------------------------------------------------------------------------------------------
def build_model(device, depth, w_bits=8, smooth=None, static=False):
    import torch
    from viditq_b200.qdiff import QuantModel
    from viditq_b200.stdit import STDiT
    model = STDiT(input_size=(T_FRAMES, 64, 64), depth=depth, hidden_size=HIDDEN, num_heads=HEADS)
    model.eval()
    wq, aq = quant_cfgs(w_bits, smooth, static=static)
    qnn = QuantModel(model, wq, aq)
    qnn.cfg_split = True
    if smooth is not None:      # synthetic calibration statistics (act_scale = |randn| + 0.5)
        g = torch.Generator().manual_seed(7)
        n_tr = len(smooth.get("timerange", [[0, 1000]]))
        for name, layer in qnn.quant_layers():
            layer.act_quantizer.act_scale = torch.randn(n_tr, 1, layer.in_features, generator=g).abs() + 0.5
            if not name.startswith("blocks."):
                layer.smooth_quant = False
    qnn.to(device)
    qnn.half()
    model.dtype = torch.float16
    qnn.set_module_name_for_quantizer(module=qnn.model)
    qnn.fp_layer_list = FP_LAYERS
    if static:
        # the activation scales come from this repo's PTQ producer on a synthetic calibration set (two timesteps, one
        # cond + uncond pair each): ptq.py:296-327 through the integer kernels, layer-by-layer schedule
        from viditq_b200 import ptq
        g = torch.Generator().manual_seed(11)
        xs = torch.randn(4, 4, T_FRAMES, 64, 64, generator=g)
        ts = torch.tensor([800.0, 800.0, 200.0, 200.0])
        cs = torch.randn(4, 1, PROMPT_LEN, 4096, generator=g).half()
        masks = torch.zeros(4, PROMPT_LEN, dtype=torch.int64)
        masks[:, :109] = 1
        ptq.run_ptq(qnn, (xs, ts, cs, masks), n_samples=1, batch_size=1, fp_layer_list=FP_LAYERS, device=device)
        return qnn, model
    qnn.init_weight_quant_params()
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    qnn.set_quant_state(True, True)
    return qnn, model


def build_pixart(device, depth):
    """PixArt-alpha 512 (64x64 latent -> 1024 tokens) under t2i/configs/quant/alpha/w8a8.yaml: dynamic per-token W8A8 +
    running-stat smooth-quant on the last block's fc2 (quirk Q17, quant_txt2img.py:297-300)."""
    import torch
    from viditq_b200.pixart import PixArtMS
    from viditq_b200.qdiff import QuantModel
    model = PixArtMS(input_size=64, depth=depth)
    model.eval()
    wq, aq = quant_cfgs(8, dict(alpha=0.3), n_spatial=1024, n_temporal=1)
    qnn = QuantModel(model, wq, aq, model_type="pixart")
    last = f"blocks.{depth - 1}.mlp.fc2"
    g = torch.Generator().manual_seed(7)
    dict(qnn.quant_layers())[last].act_quantizer.act_scale = torch.randn(1, 1, 4 * HIDDEN, generator=g).abs() + 0.5
    qnn.to(device)
    qnn.half()
    model.dtype = torch.float16
    qnn.set_module_name_for_quantizer(module=qnn.model)
    qnn.set_smooth_quant(False, False)
    qnn.init_weight_quant_params()
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    qnn.set_quant_state(True, True)
    qnn.set_layer_quant(model=qnn, module_name_list=FP_LAYERS_PIXART, quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")
    qnn.set_layer_smooth_quant(model=qnn, module_name_list=[last], smooth_quant=True, smooth_quant_running_stat=True)
    return qnn, model


def multi_gpu_selfcheck(dev, world, rank):
    """Driver-visible multi-GPU parity record (the 2-GPU pytest cases skip on a 1-GPU box): on every pair of ranks (2i, 2i+1)
    a small STDiT step (2 blocks, 16 frames of 32x32) run (a) as a cfg-branch pair, (b) frame-sharded over the pair — eagerly
    and as CUDA-graph segments — must be BIT-IDENTICAL to the same step on one GPU.  Returns the all-ranks AND."""
    import torch
    import torch.distributed as dist
    from viditq_b200 import ops, shard
    from viditq_b200.qdiff import QuantModel
    from viditq_b200.sampler import SpacedDDIM
    from viditq_b200.stdit import STDiT
    torch.manual_seed(11)
    model = STDiT(input_size=(16, 32, 32), depth=2, hidden_size=HIDDEN, num_heads=HEADS).eval()
    wq, aq = quant_cfgs(n_spatial=256)
    qnn = QuantModel(model, wq, aq)
    qnn.to(dev).half()
    model.dtype = torch.float16
    qnn.set_module_name_for_quantizer(module=qnn.model)
    qnn.fp_layer_list = FP_LAYERS
    qnn.init_weight_quant_params()
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    qnn.set_quant_state(True, True)
    qnn.set_timestep_id_for_quantlayer(999.0)
    g = torch.Generator().manual_seed(7)
    z = torch.randn(1, 4, 16, 32, 32, generator=g).to(dev)
    yc = torch.randn(1, 1, PROMPT_LEN, 4096, generator=g).to(dev)
    yu = torch.randn(1, 1, PROMPT_LEN, 4096, generator=g).to(dev)
    mask = torch.zeros(1, PROMPT_LEN, dtype=torch.int64)
    mask[0, :77] = 1
    mask = mask.to(dev)
    ddim = SpacedDDIM(num_sampling_steps=100, cfg_scale=4.0)
    t1 = torch.full((1,), 999.0, device=dev)
    coef = ddim.coefficients(ddim.num_timesteps - 1, "cpu").to(dev)
    plan1 = model.mask_select_plan(mask)
    seg1 = model.kv_segments(plan1[1], dev)
    plan2 = model.mask_select_plan(mask.repeat(2, 1))
    seg2 = model.kv_segments(plan2[1], dev)
    pair = shard.cfg_pair_groups()
    prank = rank % 2
    y2 = torch.cat([yc, yu])
    # single-GPU references
    oc = model.forward_fused(z, t1, yc, plan=plan1, segments=seg1)
    ou = model.forward_fused(z, t1, yu, plan=plan1, segments=seg1)
    ref_step = ops.cfg_ddim_step(oc, ou, z, coef, ddim.cfg_scale)
    ref_fwd = model.forward_fused(torch.cat([z, z]), t1.expand(2), y2, plan=plan2, segments=seg2, independent=True)

    def branch_step():
        mine = model.forward_fused(z, t1, yu if prank else yc, plan=plan1, segments=seg1)
        a, b = shard.exchange_cfg_branches(mine, pair)
        return ops.cfg_ddim_step(a, b, z, coef, ddim.cfg_scale)
    f0, f1 = shard.frame_slice(16, 2, prank)
    z_loc = z[:, :, f0:f1].contiguous()

    def frames_fwd():
        return model.forward_fused(torch.cat([z_loc, z_loc]), t1.expand(2), y2, plan=plan2, segments=seg2, independent=True,
                                   frames=(pair, 2, prank))
    res = {}
    res["cfg_branch_bit_identical"] = bool(torch.equal(branch_step(), ref_step))
    res["frames_bit_identical"] = bool(torch.equal(frames_fwd(), ref_fwd[:, :, f0:f1]))
    sg = shard.SegmentedGraph()
    out = sg.capture(frames_fwd)
    sg.replay()
    torch.cuda.synchronize()
    res["frames_segmented_graph_bit_identical"] = bool(torch.equal(out, ref_fwd[:, :, f0:f1]))
    sg2 = shard.SegmentedGraph()
    out2 = sg2.capture(branch_step)
    sg2.replay()
    torch.cuda.synchronize()
    res["cfg_branch_segmented_graph_bit_identical"] = bool(torch.equal(out2, ref_step))
    flags = torch.tensor([int(v) for v in res.values()], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    out = {k: bool(v) for k, v in zip(res.keys(), flags.tolist())}
    out["frames_graph_segments"], out["frames_nccl_calls"] = sg.counts()
    return out


# --------------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------------
def build_w4a8mp(dev, depth):
    """BASELINE config 4: w4a8_timestep_aware_cb.yaml (4-bit weights, timerange-aware smooth-quant, alpha 0.11 / 0.11, 20
    steps, cfg 7) + per-timestep mixed precision tables shaped like t20_weight_4_mp.yaml / t20_act_8_mp.yaml: range 19-15
    all 8 bit, the others 4 bit except the MLP layers of the first and last blocks (58 of 364 layers) at 8."""
    from viditq_b200.sampler import SpacedDDIM
    qnn, model = build_model(dev, depth, w_bits=4, smooth=dict(alpha=[0.11, 0.11], timerange=[[0, 500], [501, 1000]]))
    ddim = SpacedDDIM(num_sampling_steps=20, cfg_scale=7.0)
    names = ["model." + n for n, _ in qnn.quant_layers() if n.startswith("blocks.")]
    hot = {n for n in names if ".mlp." in n and int(n.split(".")[2]) in (0, 1, depth - 1)}
    tab4 = {n: (8 if n in hot else 4) for n in names}
    qnn.timestep_wise_mp = True
    qnn.time_mp_config_weight = {"19-15": {n: 8 for n in names}, "14-10": tab4, "9-5": tab4, "4-0": tab4,
                                 "fp_layers": {k: FP_LAYERS for k in ("19-15", "14-10", "9-5", "4-0")}}
    qnn.time_mp_config_act = {k: {n: 8 for n in names} for k in ("19-15", "14-10", "9-5", "4-0")}
    return qnn, model, ddim


def run_full_sample(args, dev, world, rank):
    """BASELINE config 3 as the reference states it: ONE complete 100-step DDIM sample of the 16x512x512 STDiT under
    w8a8_dynamic.yaml (cfg_split, cfg 4.0) through viditq_b200.sampler.GraphedSampler — host noise in, host latent out,
    every step a CUDA-graph replay.  --steps = number of samples timed (each rank its own sample; default bench K would be
    10 samples = 40 s, 2 are enough), --warmup = samples run first (the first one captures the graph)."""
    import torch
    import torch.distributed as dist
    from viditq_b200 import ops
    from viditq_b200.sampler import GraphedSampler, SpacedDDIM
    mp20 = args.workload == "sample20mp"
    if mp20:      # BASELINE config 4 as one full 20-step sample: the graphed loop re-captures per (timerange, bit range)
        qnn, model, ddim = build_w4a8mp(dev, args.depth)
    else:
        qnn, model = build_model(dev, args.depth)
        ddim = SpacedDDIM(num_sampling_steps=100, cfg_scale=4.0)
    g = torch.Generator().manual_seed(99 + rank)
    h_z = torch.randn(1, 4, T_FRAMES, 64, 64, generator=g).pin_memory()
    yc = torch.randn(1, 1, PROMPT_LEN, 4096, generator=g).to(dev)
    yu = torch.randn(1, 1, PROMPT_LEN, 4096, generator=g).to(dev)
    mask = torch.zeros(1, PROMPT_LEN, dtype=torch.int64)
    mask[0, :109] = 1
    gs = GraphedSampler(qnn, model, ddim, yc, yu, mask.to(dev), h_z.shape)
    h_out = torch.empty_like(h_z).pin_memory()
    n_samples = max(1, min(args.steps, 3))
    for _ in range(max(1, min(args.warmup, 1))):
        gs.sample(h_z.to(dev, non_blocking=True))
    torch.cuda.synchronize()
    clk = ClockSampler(dev.index or 0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clk:
        clk.t0 = time.time()
        e0.record()
        for _ in range(n_samples):
            h_out.copy_(gs.sample(h_z.to(dev, non_blocking=True)), non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        clk.t1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    if not torch.isfinite(h_out).all() or ops.check_status() != 0:
        raise SystemExit("sample100: non-finite latent")
    if rank == 0:
        sec = ms.item() / 1e3 / n_samples
        print(json.dumps({
            "metric": ("stdit_16x512x512_w4a8_mp_20step_sample_seconds" if mp20 else
                       "stdit_16x512x512_w8a8_100step_sample_seconds"), "value": sec, "unit": "s/sample", "n_gpus": world,
            "steps": n_samples, "warmup": 1, "ms_per_step": sec * 1e3 / ddim.num_timesteps, "higher_is_better": False,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": ("STDiT-XL/2 16x512x512 W4A8 timestep-aware smooth-quant + per-timestep mixed precision, ONE full "
                                    "20-step DDIM sample per rank (cfg 7.0)" if mp20 else
                                    "STDiT-XL/2 16x512x512 W8A8 per-token dynamic, ONE full 100-step DDIM sample per rank "
                                    "(cfg_split stacked, cfg 4.0)") + " through sampler.GraphedSampler: pinned host noise in, "
                                   "pinned host latent out, one CUDA-graph replay per step",
                       "graphs_captured": len(gs.graphs),
                       "samples_timed_per_gpu": n_samples, "denoise_steps_per_sec_per_gpu": ddim.num_timesteps / sec,
                       "depth": args.depth, "l2": "working set per step exceeds the 126 MB L2"},
            "e2e": {"value": sec, "unit": "s/sample", "h2d_bytes_per_step": h_z.numel() * 4 // ddim.num_timesteps,
                    "d2h_bytes_per_step": h_out.numel() * 4 // ddim.num_timesteps},
            "gpu_launches": gs.launches_per_step * ddim.num_timesteps * n_samples, "clocks": clk.summary()}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="stdit", choices=["stdit", "linear", "pixart512", "w4a8mp", "w8a8static", "sample100", "sample20mp"],
                    help="stdit: the metric (BASELINE configs 3 / 5). linear: config 1, one QuantLinear 1152->4608 through the "
                         "hook API. pixart512: config 2, one PixArt-alpha 512 solver step (CFG batch 2) under w8a8.yaml. "
                         "w4a8mp: config 4, STDiT W4A8 timestep-aware smooth-quant + per-layer mixed precision")
    ap.add_argument("--schedule", default="fused", choices=["fused", "hook", "hook-graph"],
                    help="fused: forward_fused in a CUDA graph. hook: the reference's hook API — QuantModel.forward, one "
                         "QuantLayer call per linear, eager (its .item() / mask_select host syncs are the reference's own)")
    ap.add_argument("--depth", type=int, default=DEPTH, help="debug only: fewer blocks (result is then NOT the metric)")
    ap.add_argument("--no-graph", action="store_true", help="debug: eager launches instead of a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-selfcheck", action="store_true", help="skip the multi-GPU bit-identity self-check (N >= 2)")
    ap.add_argument("--no-peak", action="store_true", help="skip the INT8 peak measurement (roofline then uses the proxy)")
    ap.add_argument("--cfg-mode", default="stacked", choices=["stacked", "split", "split2"],
                    help="cfg_split's cond / uncond forwards as one stacked launch sequence (default), as two calls, or as two "
                         "calls on two CUDA streams (split2: one branch's tail waves and HBM-bound passes may overlap the other's)")
    ap.add_argument("--attn-int8", action="store_true",
                    help="OPT-IN: spatial attention on INT8 Q/K/V (vq_attn_spatial_i8).  Leaves the reference's numerics (its "
                         "attention is fp16): the line is then NOT the headline metric and says so in config")
    ap.add_argument("--parallelism", default="samples", choices=["samples", "cfg-branch", "frames"],
                    help="samples: one sample per rank, no data-path collective (the metric, weak scaling). cfg-branch: "
                         "two ranks per sample, one CFG branch each, model outputs exchanged every step (latency / "
                         "strong scaling of one sample; needs an even --gpus). frames: ONE sample, its 16 frames sharded "
                         "over all ranks, codes all-to-all around the temporal attention of every block (strong scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.attn_int8:
        os.environ["VQ_ATTN_INT8"] = "1"      # read by stdit.FusedBlocks when the schedule is built

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: viditq_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from viditq_b200 import ops, shard
    from viditq_b200.sampler import SpacedDDIM, TimestepMixedPrecision
    torch.set_grad_enabled(False)
    pairs = args.parallelism == "cfg-branch"
    if pairs and (world < 2 or world % 2):
        raise SystemExit("--parallelism cfg-branch needs an even number of ranks")
    pair_group = shard.cfg_pair_groups() if pairs else None
    fsh = args.parallelism == "frames" and world > 1
    if (pairs or fsh) and args.workload not in ("stdit",) or (args.schedule != "fused" and
                                                               args.workload not in ("stdit", "w8a8static")):
        raise SystemExit("--parallelism / --schedule hook apply to --workload stdit")
    sample_id = rank // 2 if pairs else (0 if fsh else rank)   # ranks sharing a sample hold the same inputs and weights
    torch.manual_seed(1234 + sample_id)
    wl = args.workload
    if wl in ("sample100", "sample20mp"):
        return run_full_sample(args, dev, world, rank)
    metric, unit = METRIC, UNIT
    peaks = load_peaks()

    # ================================================================================================================
    # workload set-up: every branch defines step_device() (inputs resident), the host buffers of the e2e loop and a label
    # ================================================================================================================
    g = torch.Generator().manual_seed(99 + sample_id)
    use_graph = not args.no_graph
    if wl in ("stdit", "w4a8mp", "w8a8static"):
        if wl == "w8a8static":
            # w8a8_naive.yaml: static per-tensor activation scales, calibrated here by viditq_b200.ptq; same step otherwise
            qnn, model = build_model(dev, args.depth, static=True)
            ddim = SpacedDDIM(num_sampling_steps=100, cfg_scale=4.0)
            mp = None
            metric = "stdit_16x512x512_w8a8_static_denoise_steps_per_sec"
        elif wl == "w4a8mp":
            qnn, model, ddim = build_w4a8mp(dev, args.depth)
            mp = TimestepMixedPrecision(qnn)
            metric = "stdit_16x512x512_w4a8_mp_denoise_steps_per_sec"
        else:
            qnn, model = build_model(dev, args.depth)
            ddim = SpacedDDIM(num_sampling_steps=100, cfg_scale=4.0)
            mp = None
        h_z = torch.randn(1, 4, T_FRAMES, 64, 64, generator=g).pin_memory()
        h_yc = torch.randn(1, 1, PROMPT_LEN, 4096, generator=g).pin_memory()
        h_yu = torch.randn(1, 1, PROMPT_LEN, 4096, generator=g).pin_memory()
        mask = torch.zeros(1, PROMPT_LEN, dtype=torch.int64)
        mask[0, :109] = 1                                     # a 109-token prompt (text_embeds.pth has 101..120)
        h_t = torch.zeros(1).pin_memory()
        h_coef = torch.zeros(4).pin_memory()
        h_out = torch.empty(1, 4, T_FRAMES, 64, 64).pin_memory()
        if fsh:   # this rank's frames of the latent (it stays frame-sharded through the whole sampling loop)
            f0, f1 = shard.frame_slice(T_FRAMES)
            h_z = h_z[:, :, f0:f1].contiguous().pin_memory()
            h_out = torch.empty_like(h_z).pin_memory()
        d_z = h_z.to(dev)
        d_y = torch.cat([h_yc, h_yu]).to(dev)                 # cond | uncond captions, one stacked device buffer
        d_yc, d_yu = d_y[:1], d_y[1:]
        d_t, d_coef = torch.zeros(1, device=dev), torch.zeros(4, device=dev)
        d_mask = mask.to(dev)
        plan = model.mask_select_plan(mask.repeat(2, 1).to(dev))
        segments = model.kv_segments(plan[1], dev)
        plan1 = model.mask_select_plan(d_mask)
        segments1 = model.kv_segments(plan1[1], dev)
        sched = [(ddim.model_timestep(i), ddim.coefficients(i, "cpu")) for i in range(ddim.num_timesteps)]
        # w4a8mp: the timed steps sit in ONE mixed-precision range (step index 12: W4 with the 8-bit MLP layers) so that a
        # captured graph stays valid; stdit: steps walk down the 100-step schedule (same kernels at every step)
        fixed_i = 12 if wl == "w4a8mp" else None

        def set_step(i):
            i = fixed_i if fixed_i is not None else ddim.num_timesteps - 1 - (i % ddim.num_timesteps)
            h_t.fill_(sched[i][0])
            h_coef.copy_(sched[i][1])
            return i
        hook = args.schedule in ("hook", "hook-graph")
        hook_graph = args.schedule == "hook-graph"
        if hook and not hook_graph:
            use_graph = False
        side_streams = [torch.cuda.Stream(), torch.cuda.Stream()] if args.cfg_mode == "split2" else []

        def step_device():
            """The denoise step on device-resident inputs (iddpm forward_with_cfg + ddim_sample, cfg_split): the cond and
            uncond forwards of cfg_split run as one stacked launch sequence with un-pooled statistics (== two batch-1 calls,
            tests/test_gpu_stdit.py::test_stacked_cfg_split_equals_two_separate_forwards)."""
            if hook_graph:   # the same layer-by-layer schedule, sync-free so that a CUDA graph can hold it: the wrapped model
                # is called directly (QuantModel.forward reads t[0].item(); the timestep is set outside the graph) with
                # the mask-select plan computed once per prompt
                out_c = qnn.model(d_z, d_t, d_yc, plan=plan1)
                out_u = qnn.model(d_z, d_t, d_yu, plan=plan1)
            elif hook:    # the reference's own call sequence: QuantModel.forward twice (iddpm/__init__.py:156-157)
                out_c = qnn(d_z, d_t, d_yc, mask=d_mask)
                out_u = qnn(d_z, d_t, d_yu, mask=d_mask)
            elif fsh:
                out = model.forward_fused(torch.cat([d_z, d_z]), d_t.expand(2), d_y, plan=plan, segments=segments,
                                          independent=True, frames=(None, world, rank))
                out_c, out_u = out[:1], out[1:]
            elif pairs:   # this rank's branch only; the 2 MB outputs cross NVLink, then both ranks apply the same update
                mine = model.forward_fused(d_z, d_t, d_yu if shard.cfg_branch() else d_yc, plan=plan1, segments=segments1)
                out_c, out_u = shard.exchange_cfg_branches(mine, pair_group)
            elif args.cfg_mode == "stacked":
                out = model.forward_fused(torch.cat([d_z, d_z]), d_t.expand(2), d_y, plan=plan, segments=segments,
                                          independent=True)
                out_c, out_u = out[:1], out[1:]
            elif args.cfg_mode == "split2":
                cur = torch.cuda.current_stream()
                for st_ in side_streams:
                    st_.wait_stream(cur)
                with torch.cuda.stream(side_streams[0]):
                    out_c = model.forward_fused(d_z, d_t, d_yc, plan=plan1, segments=segments1)
                with torch.cuda.stream(side_streams[1]):
                    out_u = model.forward_fused(d_z, d_t, d_yu, plan=plan1, segments=segments1)
                for st_ in side_streams:
                    cur.wait_stream(st_)
            else:
                out_c = model.forward_fused(d_z, d_t, d_yc, plan=plan1, segments=segments1)
                out_u = model.forward_fused(d_z, d_t, d_yu, plan=plan1, segments=segments1)
            return ops.cfg_ddim_step(out_c, out_u, d_z, d_coef, ddim.cfg_scale)

        i0 = set_step(0)
        if mp is not None:
            mp.before_step(i0)
        d_t.copy_(h_t)
        d_coef.copy_(h_coef)
        qnn.set_timestep_id_for_quantlayer(float(h_t[0]))
        h_in = (h_z, h_yc, h_yu, h_t, h_coef)
        d_in = (d_z, d_yc, d_yu, d_t, d_coef)
        label = ("STDiT-XL/2 16x512x512 (T=16,S=1024 -> 16384 tokens, 28 blocks) W8A8 per-token dynamic (w8a8_dynamic.yaml), "
                 "cfg_split: cond + uncond forwards (one stacked launch sequence, un-pooled statistics == two batch-1 calls) "
                 "+ CFG + DDIM per step") if wl == "stdit" else (
                 "STDiT-XL/2 16x512x512 W8A8 with STATIC per-tensor activation scales (w8a8_naive.yaml; calibrated by "
                 "viditq_b200.ptq on a synthetic calibration set), cfg_split stacked + CFG + DDIM per step") if wl == "w8a8static" else (
                 "STDiT-XL/2 16x512x512 W4A8 (w4a8_timestep_aware_cb.yaml: 4-bit weights, timerange-aware smooth-quant) + "
                 "per-layer mixed precision (t20_*_mp.yaml shape: MLP layers of 3 blocks at 8 bit), cfg 7.0, step 12 of 20")
        if hook_graph:
            label += (" — HOOK schedule in a CUDA graph: the wrapped model's layer-by-layer forward x2 (13 QuantLayer calls per "
                      "block, torch SDPA attention), mask-select plan precomputed, timestep set outside the graph")
        elif hook:
            label += " — HOOK schedule: QuantModel.forward x2, 13 QuantLayer calls per block, torch SDPA attention, eager"
        total_linear_top = 2 * linear_ops_per_forward() / 1e12 * args.depth / DEPTH
    elif wl == "pixart512":
        qnn, model = build_pixart(dev, args.depth)
        metric, unit = "pixart_alpha_512_w8a8_solver_steps_per_sec", "steps/s"
        h_z = torch.randn(1, 4, 64, 64, generator=g).pin_memory()
        h_y = torch.randn(2, 1, PROMPT_LEN, 4096, generator=g).pin_memory()      # cond | null caption
        h_t = torch.full((2,), 500.0).pin_memory()
        h_out = torch.empty(2, 8, 64, 64, dtype=torch.float16).pin_memory()
        mask = torch.zeros(2, PROMPT_LEN, dtype=torch.int64)
        mask[:, :109] = 1
        d_z, d_y, d_t = h_z.to(dev), h_y.to(dev), h_t.to(dev)
        plan = model.mask_select_plan(mask.to(dev))
        segments = model.kv_segments(plan[1], dev)

        def set_step(i):
            return i

        def step_device():
            # dpm-solver model_fn (t2i/diffusion/model/dpm_solver.py): ONE forward of the cond | uncond batch per solver step
            return model.forward_fused(torch.cat([d_z, d_z]), d_t, d_y, plan=plan, segments=segments)
        h_in, d_in = (h_z, h_y, h_t), (d_z, d_y, d_t)
        label = ("PixArt-alpha XL/2 512x512 (64x64 latent -> 1024 tokens, 28 blocks), CFG batch 2 (M = 2048, pooled statistics), "
                 "w8a8.yaml: dynamic per-token W8A8 + running-stat smooth-quant on blocks.27.mlp.fc2 (Q17); one solver step")
        total_linear_top = pixart_linear_ops(args.depth) / 1e12
    else:   # linear: BASELINE config 1, through the hook API of one layer
        from viditq_b200.qdiff import QuantLayer
        metric, unit = "quantlinear_w8a8_1152x4608_m16384_forwards_per_sec", "forwards/s"
        lin = torch.nn.Linear(HIDDEN, 4 * HIDDEN)
        wq, aq = quant_cfgs()
        layer = QuantLayer(lin, wq, aq).to(dev).half()
        w = layer.weight.data.float()
        mn, mx = w.min(1)[0].clamp(max=0), w.max(1)[0].clamp(min=0)
        layer.weight_quantizer.delta = ((mx - mn) / 255).half().reshape(-1, 1)
        layer.weight_quantizer.zero_point = torch.round(-mn / ((mx - mn) / 255)).half().reshape(-1, 1)
        layer.weight_quantizer.init_done = layer.act_quantizer.init_done = True
        layer.set_quant_state(True, True)
        M = T_FRAMES * S_TOKENS
        h_x = torch.randn(1, M, HIDDEN, generator=g).half().pin_memory()
        h_out = torch.empty(1, M, 4 * HIDDEN, dtype=torch.float16).pin_memory()
        d_x = h_x.to(dev)

        def set_step(i):
            return i

        def step_device():
            return layer(d_x)
        h_in, d_in = (h_x,), (d_x,)
        label = "single QuantLinear W8A8 in=1152 out=4608, x [1, 16384, 1152] fp16, QuantLayer.forward (hook API) -> vq_linear_w8a8"
        total_linear_top = 2.0 * M * HIDDEN * 4 * HIDDEN / 1e12

    # ---- eager warm-up (builds prepared weights), launch count and algorithmic work of one step --------------------
    n0 = ops.launch_count()
    d_out = step_device()
    torch.cuda.synchronize()
    ops.check_status()
    with WorkMeter(ops) as meter:
        n0 = ops.launch_count()
        step_device()
        launches_per_step = ops.launch_count() - n0            # steady state (no weight prep)
    torch.cuda.synchronize()
    work = {k: tuple(v) for k, v in meter.work.items()}

    graph = seg_graph = None
    # steps with NCCL calls inside (cfg-branch / frames): CUDA-graph SEGMENTS around eagerly issued collectives
    # (shard.SegmentedGraph) — capturing the torch.distributed calls inside one whole-step graph deadlocked on the 2-GPU box
    if use_graph and (pairs or fsh):
        seg_graph = shard.SegmentedGraph()
        d_out = seg_graph.capture(step_device)
        torch.cuda.synchronize()
    elif use_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step_device()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            d_out = step_device()

    def run_step():
        if graph is not None:
            graph.replay()
            return d_out
        if seg_graph is not None:
            seg_graph.replay()
            return d_out
        return step_device()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # small workloads (PixArt step, single layer) fit the 126 MB L2: flush it between timed iterations
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if wl in ("pixart512", "linear") else None

    def timed_loop(n, body):
        """n iterations; with an L2 flush between them each iteration is bracketed by its own event pair."""
        if flush is None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                body(i)
            e1.record()
            barrier()
            return e0.elapsed_time(e1)
        evs = []
        for i in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            body(i)
            e1.record()
            evs.append((e0, e1))
        barrier()
        return sum(a.elapsed_time(b) for a, b in evs)

    # ---- value: K steps, inputs resident in HBM ---------------------------------------------------------------
    with ClockSampler(local) as clk:       # nvidia-smi streams from here on; only samples inside the timed region count
        for _ in range(args.warmup):
            run_step()
        barrier()
        clk.t0 = time.time()
        ms = timed_loop(args.steps, lambda i: run_step())
        clk.t1 = time.time()
    # ---- e2e: host buffers in, host result out, every step ----------------------------------------------------
    for _ in range(max(1, args.warmup // 2)):
        run_step()
    barrier()

    def e2e_body(i):
        set_step(i)
        for d, h in zip(d_in, h_in):
            d.copy_(h, non_blocking=True)
        out = run_step()
        h_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the caller reads the result before the next step
    ms_e2e = timed_loop(args.steps, e2e_body)
    ops.check_status()
    h2d = sum(t.numel() * t.element_size() for t in h_in)
    d2h = h_out.numel() * h_out.element_size()

    # ---- per-kernel times inside the timed schedule (CUPTI over one replay) + roofline ---------------------------------
    prof, busy_ms, span_ms = profile_kernels(run_step)
    int8_peak = None if (args.no_peak or rank != 0) else measure_int8_peak(dev)

    parity = None
    if world > 1:
        tt = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = tt.tolist()
        if world % 2 == 0 and wl == "stdit" and not args.no_selfcheck:
            try:
                parity = multi_gpu_selfcheck(dev, world, rank)
            except Exception as e:      # the record must not cost the bench line
                parity = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
    if rank == 0:
        n_samples = world // 2 if pairs else (1 if fsh else world)
        value = n_samples * args.steps / (ms * 1e-3)
        e2e_value = n_samples * args.steps / (ms_e2e * 1e-3)
        roofs = rooflines(work, prof, peaks, int8_peak, busy_ms, ms / args.steps)
        main_roof = next((r for r in roofs if r["kernel_class"] == "gemm"), None)
        roofline = None
        if main_roof is not None:
            roofline = {"bound": "tensor", "achieved": main_roof["achieved"], "peak": main_roof["peak"], "unit": "TOP/s",
                        "frac": main_roof["frac"], "traffic": main_roof["traffic"],
                        "traffic_source": main_roof["traffic_source"],
                        "algorithmic_bytes_per_launch": main_roof["algorithmic_bytes_per_launch"],
                        "kernel": "vq_gemm_w8a8_kernel / vq_linear_fused_kernel (all QuantLinear GEMMs of a step)",
                        "peak_source": main_roof["peak_source"],
                        "timing": ("share of the kernel class in a CUPTI-profiled replay of the timed CUDA graph x ms_per_step of "
                                   "the timed region" if graph is not None else
                                   "share of the kernel class in a CUPTI-profiled eager step x ms_per_step of the timed region"),
                        "gemm_ms_per_step": main_roof["ms_per_step"], "gemm_launches_per_step": main_roof["launches_per_step"],
                        "whole_step_frac": total_linear_top / (ms / args.steps * 1e-3) / main_roof["peak"],
                        "int8_peak_measured": int8_peak}
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if (pairs or fsh) else "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": label,
                       "samples_per_gpu": 0.5 if pairs else (1.0 / world if fsh else 1),
                       "parallelism": (f"cfg-branch pairs x{world // 2}: one CFG branch per rank, all_gather of the model "
                                       f"outputs (2 MB) per step" if pairs else
                                       f"frame-sharded x{world}: {T_FRAMES // world} frames per rank, all-to-all of the "
                                       f"temporal branch's u8 codes per block" if fsh else
                                       f"sample-sharded x{world} (no data-path collective)"),
                       "schedule": args.schedule,
                       "cuda_graph": (graph is not None) or (seg_graph is not None and
                                                             "%d graph segments around %d NCCL calls" % seg_graph.counts()),
                       "depth": args.depth,
                       "cfg_mode": args.cfg_mode,
                       "spatial_attention": ("INT8 Q/K/V (opt-in vq_attn_spatial_i8: NOT the reference's numerics, own tolerance)"
                                             if args.attn_int8 else "fp16 (the reference's arithmetic)"),
                       "l2": ("256 MB buffer written between timed iterations (working set fits the 126 MB L2)" if flush is not None
                              else "working set per step (0.74 GB weight codes + >1 GB activations) exceeds the 126 MB L2"),
                       "linear_TOP_per_step": total_linear_top},
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clk.summary(),
            "roofline": roofline,
            "roofline_kernels": roofs,
            "step_busy_ms_cupti": busy_ms, "step_span_ms_cupti": span_ms,
        }
        if parity is not None:
            line["multi_gpu_parity"] = parity
        if world == 1 and not args.no_cpu_baseline and wl in ("stdit", "w4a8mp"):
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
