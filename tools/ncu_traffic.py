"""Per-kernel-class DRAM traffic from an ncu capture of the bench command -> profiles/r02_ncu_traffic.json (read by bench.py).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 4000 --csv \
        --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 1 --no-graph --depth 2 --no-cpu-baseline --no-peak
    python tools/ncu_traffic.py gpurun_out/traffic.csv [--skip-first N]

The capture runs the SAME launches as the timed step (stacked cfg_split: M = 32768) at depth 2; per class the bytes are
averaged over the launches of the LAST eager step in the log (steady state, prepared weights cached).  ncu serialises
kernels and flushes nothing between them, so a kernel's inputs may still sit in L2 from its producer exactly as in the
real step; absolute times are not used.
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        rows.append(r)
    per_launch = {}          # launch id -> {name, read, write, time}
    for r in rows:
        lid = r.get("ID")
        d = per_launch.setdefault(lid, {"name": r.get("Kernel Name", ""), "read": 0.0, "write": 0.0, "ns": 0.0})
        val = float(str(r.get("Metric Value", "0")).replace(",", "") or 0)
        unit = r.get("Metric Unit", "")
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}.get(unit, 1.0)
        name = r.get("Metric Name", "")
        if name == "dram__bytes_read.sum":
            d["read"] = val * mult
        elif name == "dram__bytes_write.sum":
            d["write"] = val * mult
        elif name == "gpu__time_duration.sum":
            d["ns"] = val * mult
    launches = [per_launch[k] for k in sorted(per_launch, key=lambda s: int(s))]
    # keep the last step only: everything after the last-but-one vq_cfg_ddim_kernel
    ends = [i for i, l in enumerate(launches) if "vq_cfg_ddim_kernel" in l["name"]]
    if len(ends) >= 2:
        launches = launches[ends[-2] + 1: ends[-1] + 1]
    out = {}
    for cls, pat, _ in bench.KERNEL_CLASSES:
        sel = [l for l in launches if re.search(pat, l["name"])]
        if not sel:
            continue
        out[cls] = {"launches": len(sel),
                    "dram_bytes_per_launch": sum(l["read"] + l["write"] for l in sel) / len(sel),
                    "dram_read_bytes_per_launch": sum(l["read"] for l in sel) / len(sel),
                    "dram_write_bytes_per_launch": sum(l["write"] for l in sel) / len(sel),
                    "avg_us_under_ncu": sum(l["ns"] for l in sel) / len(sel) / 1e3,
                    "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum, {os.path.basename(path)}: last eager step of "
                              f"`bench.py --depth 2 --no-graph` (launched shapes, M = 32768), mean over {len(sel)} launches"}
    dst = bench.TRAFFIC_FILE
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))
    # per-kernel-name table for profiles/
    agg = {}
    for l in launches:
        key = re.sub(r"\(.*", "", l["name"])[:90]
        a = agg.setdefault(key, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += l["read"]
        a[2] += l["write"]
        a[3] += l["ns"]
    print("| kernel | launches | avg us (ncu) | avg DRAM read MB | avg DRAM write MB |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][3]):
        print(f"| `{k}` | {a[0]} | {a[3] / a[0] / 1e3:.1f} | {a[1] / a[0] / 1e6:.1f} | {a[2] / a[0] / 1e6:.1f} |")


if __name__ == "__main__":
    main()
