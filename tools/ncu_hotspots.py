"""Top stall hot spots of an ncu report (SASS view): python tools/ncu_hotspots.py report.ncu-rep [kernel-regex] [N]."""
import csv, subprocess, sys, io
rep = sys.argv[1]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several kernels may be concatenated: split on "Kernel Name"
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    if len(sys.argv) > 2 and sys.argv[2] not in b["name"]:
        continue
    hdr, body = b["rows"][0], b["rows"][1:]
    ci, si, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(float(r[ci] or 0) for r in body) or 1.0
    print("==", b["name"][:100], "samples", int(tot), "instructions", int(sum(float(r[ie] or 0) for r in body)))
    for idx, r in sorted(enumerate(body), key=lambda t: -float(t[1][ci] or 0))[:n]:
        reasons = sorted(((float(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        why = ", ".join(f"{nm} {100*v/max(float(r[ci]),1):.0f}%" for v, nm in reasons if v > 0)
        print(f"{100*float(r[ci])/tot:5.1f}%  #{idx:4d} {r[si].strip()[:70]:70s} [{why}]")
