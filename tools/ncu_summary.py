"""Text summary of an ncu report for profiles/: key metrics, stall breakdown, top SASS hot spots.

    python tools/ncu_summary.py report.ncu-rep [kernel-substring] > profiles/rN/xyz_summary.txt
"""
import csv, io, subprocess, sys

rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
print(f"# ncu summary of {rep.split('/')[-1]} (ncu --set full --clock-control none; per-launch values, cold caches, serialised)")
for r in rows[2:]:
    if flt and flt not in r[ki]:
        continue
    print(f"\n== {r[ki][:120]}")
    for k in WANT:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:70s} {r[i]} {units[i]}")
    st = {k[len('smsp__pcsamp_warps_issue_stalled_'):]: float(v or 0) for k, v in zip(hdr, r)
          if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")}
    tot = sum(st.values()) or 1.0
    print("  warp-state samples: " + ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
seen = set()
for b in blocks:
    if (flt and flt not in b["name"]) or b["name"] in seen or len(b["rows"]) < 2:
        continue
    seen.add(b["name"])
    h, body = b["rows"][0], b["rows"][1:]
    ci, si = h.index("# Samples"), h.index("Source")
    cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(float(r[ci] or 0) for r in body) or 1.0
    print(f"\n-- top SASS hot spots of {b['name'][:90]} (share of warp-state samples, dominant reasons)")
    for idx, r in sorted(enumerate(body), key=lambda t: -float(t[1][ci] or 0))[:14]:
        why = sorted(((float(r[i] or 0), h[i][6:]) for i in cols), reverse=True)[:2]
        print(f"  {100*float(r[ci])/tot:5.1f}%  #{idx:5d} {r[si].strip()[:64]:64s} [" +
              ", ".join(f"{n} {100*v/max(float(r[ci]),1):.0f}%" for v, n in why if v > 0) + "]")
