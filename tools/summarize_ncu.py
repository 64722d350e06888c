"""Summarise gpurun_out/ ncu artefacts into small text files under profiles/ (the .ncu-rep files themselves are scratch).

  python tools/summarize_ncu.py r01            -> profiles/r01_launches.md, profiles/r01_<report>.md ...
"""
import collections
import csv
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(tag):
    path = os.path.join(SRC, "launches.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) > vi and r[vi]:
            k = r[ki].split("(")[0][:90]
            agg[k][0] += 1
            agg[k][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list of `bench.py --steps 1 --warmup 3 --no-graph --depth 2` "
                f"(gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n\n")
        f.write(f"{len(data)} launches, {tot / 1e6:.2f} ms total\n\n| share | launches | avg us | kernel |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"| {v[1] / tot * 100:.2f}% | {v[0]} | {v[1] / v[0] / 1e3:.1f} | `{k}` |\n")


def report(tag, rep):
    name = os.path.splitext(os.path.basename(rep))[0]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return
    hdr = rows[0]
    with open(os.path.join(OUT, f"{tag}_{name}.md"), "w") as f:
        f.write(f"# {tag}: ncu --set full --clock-control none, {name}\n")
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, rows[1]))
            f.write(f"\n## `{d.get('Kernel Name', '?')[:110]}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {u.get(k, '')} |\n")
            for k in hdr:
                if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                    try:
                        if float(d[k]) >= 0.3:
                            f.write(f"| stall {k.split('issue_stalled_')[1].split('_per_issue')[0]} (per issue) | {float(d[k]):.2f} | |\n")
                    except ValueError:
                        pass
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))[2:]
    try:
        top = sorted((r for r in srows if len(r) > 5 and r[2].isdigit()), key=lambda r: -int(r[2]))[:15]
        with open(os.path.join(OUT, f"{tag}_{name}.md"), "a") as f:
            f.write("\n### top stall-sample instructions (first kernel)\n\n| samples | executed | SASS |\n|---|---|---|\n")
            for r in top:
                f.write(f"| {r[2]} | {r[5]} | `{r[1].strip()[:100]}` |\n")
    except Exception:
        pass


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    for rep in sorted(glob.glob(os.path.join(SRC, "*.ncu-rep"))):
        report(tag, rep)
    # session logs written beside the reports: only those of THIS session (newer than the oldest report minus an hour) —
    # gpurun_out/ accumulates across calls, an old log must not be re-filed under a new tag
    reps = glob.glob(os.path.join(SRC, "*.ncu-rep"))
    fresh = (min(os.path.getmtime(r) for r in reps) - 3600) if reps else float("inf")
    for log in ("selftest.log", "attn_selftest.log", "prof_kernels.log", "bench.log", "timeline.log", "smi.txt"):
        p = os.path.join(SRC, log)
        if os.path.exists(p) and os.path.getmtime(p) >= fresh:
            with open(p) as fi, open(os.path.join(OUT, f"{tag}_{log}"), "w") as fo:
                fo.write(fi.read())


if __name__ == "__main__":
    main()
