"""Instructions executed and stall samples per SOURCE LINE of a kernel in an ncu report (needs -lineinfo + --import-source on).

    python tools/ncu_lines.py report.ncu-rep [top]
"""
import csv, io, subprocess, sys
from collections import defaultdict

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = defaultdict(lambda: [0.0, 0.0, ""])
path, hdr = "", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        li, ii, si, ti = 0, hdr.index("Instructions Executed"), hdr.index("# Samples"), 1
    elif hdr and r[0].isdigit() and len(r) > ii:
        k = (path, int(r[0]))
        num = lambda v: float(v) if v not in ("", "-") else 0.0
        agg[k][0] += num(r[ii])
        agg[k][1] += num(r[si])
        agg[k][2] = r[ti].strip()[:90]
ti_, ts_ = sum(v[0] for v in agg.values()) or 1, sum(v[1] for v in agg.values()) or 1
print(f"total warp instructions {ti_:.3g}, samples {ts_:.0f}")
for (p, l), (ins, smp, txt) in sorted(agg.items(), key=lambda kv: -kv[1][int(__import__("os").environ.get("BY_SAMPLES", "0"))])[:top]:
    print(f"{100 * ins / ti_:5.1f}% inst {100 * smp / ts_:5.1f}% smp  {p}:{l:<4d} {txt}")
