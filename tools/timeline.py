"""In-graph kernel timeline of one denoise step (CUPTI through torch.profiler): per-kernel totals, busy time and gaps.

    python tools/timeline.py [--depth 4] [--no-graph]

ncu serialises kernels with cold caches; this is the complement: the durations and the idle gaps as they are inside the
replayed CUDA graph.  Not a bench number (CUPTI adds a little per-kernel overhead) — it ranks where the step goes.
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=4)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--hook", action="store_true", help="the layer-by-layer (hook API) schedule instead of forward_fused")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    torch.set_grad_enabled(False)
    from viditq_b200 import ops
    from viditq_b200.sampler import SpacedDDIM
    qnn, model = bench.build_model(dev, args.depth)
    ddim = SpacedDDIM(num_sampling_steps=100, cfg_scale=4.0)
    g = torch.Generator().manual_seed(99)
    d_z = torch.randn(1, 4, bench.T_FRAMES, 64, 64, generator=g).to(dev)
    d_yc = torch.randn(1, 1, bench.PROMPT_LEN, 4096, generator=g).to(dev)
    d_yu = torch.randn(1, 1, bench.PROMPT_LEN, 4096, generator=g).to(dev)
    mask = torch.zeros(1, bench.PROMPT_LEN, dtype=torch.int64)
    mask[0, :109] = 1
    d_t = torch.full((1,), float(ddim.model_timestep(ddim.num_timesteps - 1)), device=dev)
    d_coef = ddim.coefficients(ddim.num_timesteps - 1, "cpu").to(dev)
    plan = model.mask_select_plan(mask.repeat(2, 1).to(dev))
    segments = model.kv_segments(plan[1], dev)
    qnn.set_timestep_id_for_quantlayer(float(d_t[0]))

    plan1 = model.mask_select_plan(mask.to(dev))

    def step():
        if args.hook:
            oc = model(d_z, d_t, d_yc, plan=plan1)
            ou = model(d_z, d_t, d_yu, plan=plan1)
            return ops.cfg_ddim_step(oc, ou, d_z, d_coef, ddim.cfg_scale)
        o = model.forward_fused(torch.cat([d_z, d_z]), d_t.expand(2), torch.cat([d_yc, d_yu]), plan=plan,
                                segments=segments, independent=True)
        return ops.cfg_ddim_step(o[:1], o[1:], d_z, d_coef, ddim.cfg_scale)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    graph = None
    if not args.no_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
    run = graph.replay if graph is not None else step
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    plain_ms = e0.elapsed_time(e1)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        run()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    ks = [k for k in ks if not k[2].startswith("Memcpy") or True]
    if not ks:
        print("no device events captured")
        return
    span = (ks[-1][1] - ks[0][0]) / 1e3
    busy = sum(e - s for s, e, _ in ks) / 1e3
    gaps = [(ks[i + 1][0] - ks[i][1]) / 1e3 for i in range(len(ks) - 1)]
    gap_total = sum(g for g in gaps if g > 0)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for s, e, n in ks:
        agg[n][0] += 1
        agg[n][1] += (e - s) / 1e3
    print(f"# depth={args.depth} graph={graph is not None}: step {plain_ms:.3f} ms un-profiled; profiled span {span:.3f} ms, "
          f"{len(ks)} device events, busy {busy:.3f} ms, gaps {gap_total:.3f} ms "
          f"(mean {1e3 * gap_total / max(1, len(gaps)):.2f} us)")
    print("| share | launches | avg us | kernel |\n|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| {100 * t / span:5.2f}% | {c} | {1e3 * t / c:.1f} | `{n[:110]}` |")


if __name__ == "__main__":
    main()
