#!/usr/bin/env python
"""bench.py — STDiT 16x512x512 W8A8 denoise-steps/sec on N B200s (BASELINE.json metric), one JSON line.

A "step" is one full ViDiT-Q denoising step of one sample: cfg_split => two STDiT-XL/2 forwards (cond / uncond, 28
blocks, 16384 tokens) on the fused sm_100a kernels + CFG combine + DDIM update.  Synthetic latents / text embeddings,
seeded random-init weights, min-max weight scales.  One process per GPU; ranks hold independent samples (weak scaling,
no data-path collective — SURVEY.md §8e); value = steps of all ranks / max-over-ranks device time.

  python bench.py [--gpus N --steps K --warmup W]          our arm
  python bench.py --impl reference ...                      the reference's simulated-quant path (CPU, oracle port)
"""
import argparse
import json
import os
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


# dram__bytes_read.sum + dram__bytes_write.sum of ONE vq_gemm_w8a8_kernel launch at M = 16384, from the `ncu --set full`
# captures of the four block shapes (profiles/r01_s22_gemm_16384_<N>_<K>_<epi>.md), in MB keyed by (N, K).  Below the
# algorithmic bytes (136 / 96 / 175 / 156 MB) because inputs written by the previous kernel still sit in the 126 MB L2
# and part of the output is still there when the kernel ends.
NCU_GEMM_DRAM_MB = {(3 * HIDDEN, HIDDEN): 86.4, (HIDDEN, HIDDEN): 64.3, (4 * HIDDEN, HIDDEN): 126.6,
                    (HIDDEN, 4 * HIDDEN): 226.3}


def gemm_dram_bytes_per_step(depth):
    """Per denoise step (two forwards): 2 q|k|v, 4 hidden->hidden, fc1, fc2 GEMMs per block (kv_linear: < 1 MB, ignored)."""
    per_block = (2 * NCU_GEMM_DRAM_MB[(3 * HIDDEN, HIDDEN)] + 4 * NCU_GEMM_DRAM_MB[(HIDDEN, HIDDEN)]
                 + NCU_GEMM_DRAM_MB[(4 * HIDDEN, HIDDEN)] + NCU_GEMM_DRAM_MB[(HIDDEN, 4 * HIDDEN)])
    return 2 * depth * per_block * 1e6


def linear_ops_per_forward(n_tok=T_FRAMES * S_TOKENS, L=PROMPT_LEN):
    C = HIDDEN
    per_block = 10 * n_tok * C * C + 2 * n_tok * C * 4 * C + L * C * 2 * C
    return 2 * per_block * DEPTH


class Cfg(dict):
    __getattr__ = dict.get


def quant_cfgs():
    sq = Cfg(enable=False, channel_wise_scale_type="momentum_act_max", momentum=0.95, alpha=0.625)
    wq = Cfg(n_bits=8, per_group="channel", channel_dim=0, scale_method="min_max", round_mode="nearest",
             mixed_precision=[4, 6, 8])
    aq = Cfg(n_bits=8, per_group="token", scale_method="min_max", round_mode="nearest_ste", running_stat=False,
             dynamic=True, sym=False, n_spatial_token=S_TOKENS, n_temporal_token=T_FRAMES, n_prompt=PROMPT_LEN,
             smooth_quant=sq)
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
# CPU arm: the reference's simulated-quant path, oracle port (the Python reference cannot travel to the GPU box)
# --------------------------------------------------------------------------------------------------------------------
def cpu_block_sample(frames=4, seed=0):
    """Time one W8A8 STDiT block of the oracle on `frames` of the 16 frames (frames*1024 tokens). Returns seconds."""
    import numpy as np
    from oracle import stdit_oracle as SO
    P = SO.make_block_params(seed)
    rng = np.random.default_rng(seed + 1)
    n = frames * S_TOKENS
    x = rng.standard_normal((1, n, HIDDEN)).astype(np.float16)
    y = rng.standard_normal((1, PROMPT_LEN, HIDDEN)).astype(np.float16)
    t0 = (rng.standard_normal((1, 6 * HIDDEN)) * 0.1).astype(np.float16)
    t = time.perf_counter()
    SO.stdit_block(x, y, t0, P, frames, S_TOKENS, [PROMPT_LEN])
    return time.perf_counter() - t


def cpu_steps_per_sec(block_seconds, frames):
    # one step = 2 forwards x 28 blocks; per-token work scales with the frame count (spatial attention is per frame)
    return 1.0 / (block_seconds * (T_FRAMES / frames) * DEPTH * 2)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = 2
    for _ in range(args.warmup):
        cpu_block_sample(frames)
    times = [cpu_block_sample(frames) for _ in range(max(1, args.steps))]
    sec = sum(times) / len(times)
    val = cpu_steps_per_sec(sec, frames)
    cores = os.cpu_count()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / val, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "STDiT-XL/2 16x512x512 W8A8 (w8a8_dynamic), cfg_split, DDIM step",
                       "note": "reference = simulated-quant CPU path, numpy oracle port (Python reference cannot "
                               "travel); each step times one block on 2 of 16 frames and extrapolates x8 x28 x2"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"1 STDiT block, {frames}/16 frames ({frames * S_TOKENS} tokens), "
                                       f"{sec:.2f} s, extrapolated x{T_FRAMES // frames} frames x28 blocks x2 CFG"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------------
def build_model(device, depth):
    import torch
    from viditq_b200.qdiff import QuantModel
    from viditq_b200.stdit import STDiT
    model = STDiT(input_size=(T_FRAMES, 64, 64), depth=depth, hidden_size=HIDDEN, num_heads=HEADS)
    model.eval()
    wq, aq = quant_cfgs()
    qnn = QuantModel(model, wq, aq)
    qnn.cfg_split = True
    qnn.to(device)
    qnn.half()
    model.dtype = torch.float16
    qnn.set_module_name_for_quantizer(module=qnn.model)
    qnn.fp_layer_list = FP_LAYERS
    qnn.init_weight_quant_params()
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    qnn.set_quant_state(True, True)
    return qnn, model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--depth", type=int, default=DEPTH, help="debug only: fewer blocks (result is then NOT the metric)")
    ap.add_argument("--no-graph", action="store_true", help="debug: eager launches instead of a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cfg-mode", default="stacked", choices=["stacked", "split"],
                    help="cfg_split's cond / uncond forwards as one stacked launch sequence (default) or two calls")
    ap.add_argument("--parallelism", default="samples", choices=["samples", "cfg-branch", "frames"],
                    help="samples: one sample per rank, no data-path collective (the metric, weak scaling). cfg-branch: "
                         "two ranks per sample, one CFG branch each, model outputs exchanged every step (latency / "
                         "strong scaling of one sample; needs an even --gpus). frames: ONE sample, its 16 frames sharded "
                         "over all ranks, codes all-to-all around the temporal attention of every block (strong scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

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
    from viditq_b200.sampler import SpacedDDIM
    pairs = args.parallelism == "cfg-branch"
    if pairs and (world < 2 or world % 2):
        raise SystemExit("--parallelism cfg-branch needs an even number of ranks")
    pair_group = shard.cfg_pair_groups() if pairs else None
    fsh = args.parallelism == "frames" and world > 1
    sample_id = rank // 2 if pairs else (0 if fsh else rank)   # ranks sharing a sample hold the same inputs and weights
    torch.manual_seed(1234 + sample_id)
    torch.set_grad_enabled(False)
    qnn, model = build_model(dev, args.depth)
    ddim = SpacedDDIM(num_sampling_steps=100, cfg_scale=4.0)

    # host (pinned) inputs of one step of this rank's sample; static device buffers the graph reads
    g = torch.Generator().manual_seed(99 + sample_id)
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
    plan = model.mask_select_plan(mask.repeat(2, 1).to(dev))
    segments = model.kv_segments(plan[1], dev)
    plan1 = model.mask_select_plan(mask.to(dev))
    segments1 = model.kv_segments(plan1[1], dev)
    sched = [(ddim.model_timestep(i), ddim.coefficients(i, "cpu")) for i in range(ddim.num_timesteps)]

    def set_step(i):
        i = ddim.num_timesteps - 1 - (i % ddim.num_timesteps)
        h_t.fill_(sched[i][0])
        h_coef.copy_(sched[i][1])
        return i

    def step_device():
        """The denoise step on device-resident inputs (iddpm forward_with_cfg + ddim_sample, cfg_split): the cond and
        uncond forwards of cfg_split run as one stacked launch sequence with un-pooled statistics (== two batch-1 calls,
        tests/test_gpu_stdit.py::test_stacked_cfg_split_equals_two_separate_forwards)."""
        if fsh:
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
        else:
            out_c = model.forward_fused(d_z, d_t, d_yc, plan=plan1, segments=segments1)
            out_u = model.forward_fused(d_z, d_t, d_yu, plan=plan1, segments=segments1)
        return ops.cfg_ddim_step(out_c, out_u, d_z, d_coef, ddim.cfg_scale)

    set_step(0)
    d_t.copy_(h_t)
    d_coef.copy_(h_coef)
    qnn.set_timestep_id_for_quantlayer(float(h_t[0]))
    n0 = ops.launch_count()
    d_out = step_device()                                  # eager warm-up: builds prepared weights, sets attributes
    torch.cuda.synchronize()
    ops.check_status()
    launches_per_step = ops.launch_count() - n0
    n0 = ops.launch_count()
    step_device()
    launches_per_step = ops.launch_count() - n0            # steady state (no weight prep)
    torch.cuda.synchronize()

    # ---- roofline of the dominant kernel (vq_gemm_w8a8): instrumented eager pass (warm allocator, before graph
    # capture), CUDA events on the launching stream around every GEMM launch of one full step ---------------------
    gemm_events, orig = [], ops.gemm_w8a8

    def timed_gemm(a, w, *aa, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = orig(a, w, *aa, **kw)
        e.record()
        gemm_events.append((s, e, 2.0 * a.G * a.rows * w.N * w.K))
        return r
    ops.gemm_w8a8 = timed_gemm
    step_device()
    torch.cuda.synchronize()
    ops.gemm_w8a8 = orig
    gemm_ms = sum(s.elapsed_time(e) for s, e, _ in gemm_events)
    gemm_ops = sum(o for _, _, o in gemm_events)


    graph = None
    # steps with NCCL calls inside (cfg-branch / frames) are launched eagerly: capturing the torch.distributed calls in
    # the step graph deadlocked on the 2-GPU box (both ranks hung in capture; measured once, not pursued)
    if not args.no_graph and not (pairs or fsh):
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
        return step_device()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: K steps, inputs resident in HBM ---------------------------------------------------------------
    with ClockSampler(local) as clk:       # nvidia-smi streams from here on; only samples inside the timed region count
        for _ in range(args.warmup):
            run_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clk.t0 = time.time()
        e0.record()
        for _ in range(args.steps):
            run_step()
        e1.record()
        barrier()
        clk.t1 = time.time()
    ms = e0.elapsed_time(e1)
    # ---- e2e: host buffers in, host result out, every step ----------------------------------------------------
    for _ in range(max(1, args.warmup // 2)):
        run_step()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(args.steps):
        set_step(i)
        d_z.copy_(h_z, non_blocking=True)
        d_yc.copy_(h_yc, non_blocking=True)
        d_yu.copy_(h_yu, non_blocking=True)
        d_t.copy_(h_t, non_blocking=True)
        d_coef.copy_(h_coef, non_blocking=True)
        out = run_step()
        h_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the caller reads the new latent before the next step
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    ops.check_status()
    h2d = sum(t.numel() * t.element_size() for t in (h_z, h_yc, h_yu, h_t, h_coef))
    d2h = h_out.numel() * h_out.element_size()

    if world > 1:
        tt = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = tt.tolist()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16_sus = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_tops = 2.0 * bf16_sus
    achieved = gemm_ops / (gemm_ms * 1e-3) / 1e12
    if rank == 0:
        n_samples = world // 2 if pairs else (1 if fsh else world)
        value = n_samples * args.steps / (ms * 1e-3)
        e2e_value = n_samples * args.steps / (ms_e2e * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if (pairs or fsh) else "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "STDiT-XL/2 16x512x512 (T=16,S=1024 -> 16384 tokens, 28 blocks) W8A8 per-token "
                                   "dynamic (w8a8_dynamic.yaml), cfg_split: cond + uncond forwards (one stacked launch sequence, "
                                   "un-pooled statistics == two batch-1 calls) + CFG + DDIM per step",
                       "samples_per_gpu": 0.5 if pairs else (1.0 / world if fsh else 1),
                       "parallelism": (f"cfg-branch pairs x{world // 2}: one CFG branch per rank, all_gather of the model "
                                       f"outputs (2 MB) per step" if pairs else
                                       f"frame-sharded x{world}: {T_FRAMES // world} frames per rank, all-to-all of the "
                                       f"temporal branch's u8 codes per block" if fsh else
                                       f"sample-sharded x{world} (no data-path collective)"),
                       "cuda_graph": graph is not None, "depth": args.depth, "cfg_mode": args.cfg_mode,
                       "l2": "working set per step (0.74 GB weight codes + >1 GB activations) exceeds the 126 MB L2",
                       "linear_TOP_per_step": 2 * linear_ops_per_forward() / 1e12 * args.depth / DEPTH},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clk.summary(),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tops, "unit": "TOP/s",
                         "frac": achieved / peak_tops,
                         "traffic": gemm_dram_bytes_per_step(args.depth) / max(1, len(gemm_events)),
                         "traffic_unit": "bytes per launch (ncu dram read+write of the 4 block shapes at M=16384, "
                                         "profiles/r01_s22_gemm_*.md, averaged over this step's launches)",
                         "algorithmic_bytes_per_launch": 2 * args.depth * (2 * 136e6 + 3 * 95.6e6 + 57.9e6 + 175e6 + 156e6)
                                                         / max(1, len(gemm_events)),
                         "kernel": "vq_gemm_w8a8_kernel (all QuantLinear GEMMs of a step)",
                         "peak_source": "2 x bf16_tflops_sustained of MEASURED_PEAKS.json (INT8 dense = 2x bf16 on "
                                        "B200; no measured INT8 figure exists)" if peaks else "2 x 1400 fallback",
                         "gemm_ms_per_step": gemm_ms, "gemm_launches_per_step": len(gemm_events)},
        }
        if world == 1 and not args.no_cpu_baseline:
            frames = 2
            sec = cpu_block_sample(frames)
            line["cpu_baseline"] = {"value": cpu_steps_per_sec(sec, frames), "unit": UNIT, "cores": os.cpu_count(),
                                    "kind": "port",
                                    "sample": f"oracle (numpy) W8A8 fake-quant STDiT block, {frames}/16 frames, "
                                              f"{sec:.2f} s, extrapolated x{T_FRAMES // frames} x28 blocks x2 CFG"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
