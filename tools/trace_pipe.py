"""Per-tile timeline of two warps of CTA 0 through the fused PCG engine's tile pipeline (a -DDPCG_PIPE_TRACE build selected
with DPCG_LIB): where does a tile's period go - waiting for the window, for the stage's bytes, in the row loops, or between
tiles (element-wise updates, reductions, row extents)?"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import argparse, numpy as np, torch
import bench
import deeppreconditioning_b200 as dp
from deeppreconditioning_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--systems", type=int, default=64)
ap.add_argument("--max-iter", type=int, default=60)
ap.add_argument("--unpacked", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
args = argparse.Namespace(side=316, net="net")
net = bench.make_net(args, dev)
systems, _ = bench.build_chunk(args, list(range(a.systems)), net, dev, keep_host=False)
batch = dp.PcgBatch(systems, 1e-8, a.max_iter, pack=not a.unpacked)
batch.solve(); torch.cuda.synchronize()
cap = 1 << 15
buf = np.zeros(2 * cap, np.uint64)
_lib.check(_lib.lib().dp_debug_pipe_trace(buf.ctypes.data, cap, 0), "dp_debug_pipe_trace")
names = {(0, 1): "wait window", (0, 2): "wait bytes (1st item)", (1, 2): "wait bytes (1st item)", (2, 3): "row loop", (3, 2): "release + wait bytes (next item)",
         (3, 4): "release (+ window)", (4, 5): "tile tail: stores, reduce", (5, 0): "next tile head: loads, scalars", (4, 0): "between tiles", (5, 5): "tile without stream"}
phases = {1: "A", 2: "APPLY1", 3: "APPLY2"}
ISSUE = (6, 7)  # matrix items issued by thread 0: 6 = at the moment it needs the item itself, 7 = ahead of time
for w, who in ((0, "warp 0 (thread 0 = matrix producer)"), (1, "warp 5")):
    t = buf[w * cap:(w + 1) * cap]
    t = t[t != 0]
    lab, clk = (t >> np.uint64(48)).astype(np.int64), (t & np.uint64(0xFFFFFFFFFFFF)).astype(np.int64)
    lab, clk = lab[len(lab) // 3:], clk[len(clk) // 3:]  # skip the start-up
    if w == 0:
        pass
    if os.environ.get("RAW"):
        start = int(np.argmax(lab // 8 == 1))
        print(f"  raw marks of {who}, phase A onwards (phase.point, ns since the previous mark; point 6 = producer duty of acquire() done):")
        print("   " + " ".join(f"{l // 8}.{l % 8}:{dt}" for l, dt in zip(lab[start + 1:start + 80], np.diff(clk[start:start + 80]))))
        start = int(np.argmax(lab // 8 == 2))
        print("   " + " ".join(f"{l // 8}.{l % 8}:{dt}" for l, dt in zip(lab[start + 1:start + 80], np.diff(clk[start:start + 80]))))
    keep = ~np.isin(lab % 8, ISSUE)
    lab, clk = lab[keep], clk[keep]
    d = np.diff(clk)
    print(f"{who}:")
    for ph, pname in phases.items():
        stats, total = {}, 0
        for l0, l1, dd in zip(lab[:-1], lab[1:], d):
            if l0 // 8 == ph and l1 // 8 == ph:  # (the hop across the grid barrier is not a tile's time)
                stats.setdefault((int(l0 % 8), int(l1 % 8)), []).append(dd)
                total += dd
        ntiles = int((lab == 8 * ph + 5).sum())
        print(f"  phase {pname}: {ntiles} tiles, {total / max(ntiles, 1):.0f} ns per tile")
        for key, v in sorted(stats.items(), key=lambda kv: -np.sum(kv[1])):
            print(f"     {names.get(key, str(key)):36s} {key}: mean {np.mean(v):7.0f} ns  p90 {np.percentile(v, 90):7.0f}  n={len(v):5d}  per tile {np.sum(v) / max(ntiles, 1):6.0f} ns")
