"""Phase timeline of CTA 0 inside a batched fused solve (DPCG_TRACE=1): how long does it wait at the grid barriers?"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ["DPCG_TRACE"] = "1"
import argparse, numpy as np, torch
import bench
from deeppreconditioning_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--systems", type=int, default=64)
ap.add_argument("--max-iter", type=int, default=120)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
bench.MAX_ITER = a.max_iter
args = argparse.Namespace(systems_per_gpu=a.systems, side=316, net="net")
mine, host = bench.build_host_systems(args, 0, 1, dev)
batch = bench.device_batch(host, dev)
for _ in range(2):
    batch.reset(); batch.solve(); torch.cuda.synchronize()
cap = 4000
out = np.zeros(2 * cap, np.int64)
_lib.check(_lib.lib().dp_debug_pcg_trace(_lib.ptr(batch.ws), len(host), out.ctypes.data, cap))
lab, clk = out[0::2], out[1::2]
k = int((clk != 0).sum())
lab, clk = lab[40:k], clk[40:k]
d = np.diff(clk) / 1.9e3
names = {1: "A", 2: "APPLY1", 3: "APPLY2"}
stats = {}
for l0, l1, dd in zip(lab[:-1], lab[1:], d):
    stats.setdefault((int(l0), int(l1)), []).append(dd)
tot = sum(np.mean(v) for v in stats.values())
for (l0, l1), v in sorted(stats.items()):
    kind = "barrier wait" if l0 % 8 == 1 and l1 % 8 == 2 else "work"
    print(f"   {names.get(l0//8, l0//8)}.{l0%8} -> {names.get(l1//8, l1//8)}.{l1%8}: {np.mean(v):8.2f} us ({100*np.mean(v)/tot:4.1f} %)  {kind}")
