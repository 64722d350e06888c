"""Top stalled SASS instructions + stall-reason totals of one .ncu-rep (source page).  usage: ncu_stalls.py rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[1], rows[2:]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp] or 0) for r in data)
agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stall}
print("total samples", tot)
print(", ".join(f"{k[6:]} {v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:top_n]:
    reasons = sorted([(int(r[i] or 0), hdr[i][6:]) for i in stall], reverse=True)[:2]
    print(f"{r[isamp]:>6} {r[iex]:>8}  {r[isrc][:64]:64s} {reasons}")
