"""A/B of the fused PCG engine on the packed copies (6 B per entry) against the fp64 / int32 stream (12 B per entry): same
systems, same launch shape; prints ms per launch, solves/s, algorithmic GB/s of each format and checks the bits.
With DPCG_TRACE=1 also the phase timeline of CTA 0 of both runs."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import argparse, numpy as np, torch
import bench
import deeppreconditioning_b200 as dp
from deeppreconditioning_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--systems", type=int, default=64)
ap.add_argument("--max-iter", type=int, default=20000)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--packed-only", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
bench.MAX_ITER = a.max_iter
args = argparse.Namespace(side=316, net="net")
net = bench.make_net(args, dev)
systems, _ = bench.build_chunk(args, list(range(a.systems)), net, dev, keep_host=False)
peak = bench.peaks()[0]
trace = os.environ.get("DPCG_TRACE", "0")[:1] == "1"
names = {1: "A", 2: "APPLY1", 3: "APPLY2"}
out = {}
for tag, pack in (("fp64/int32 12 B", False), ("packed 6 B", True))[1 if a.packed_only else 0:]:
    batch = dp.PcgBatch(systems, 1e-8, a.max_iter, pack=pack)
    ms = []
    for _ in range(1 + a.reps):
        batch.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); batch.solve(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    res = batch.results()
    its = [r.iterations for r in res]
    eb = 6 if pack else 12
    nbytes = sum(bench.iter_bytes(e["A"].n, e["A"].nnz, e["M"].L.nnz, eb) * i for e, i in zip(batch.entries, its))
    best = min(ms[1:])
    print(f"{tag:16s}: {best:9.2f} ms per launch ({a.systems} systems, {sum(its)} iterations, {its[:3]}...), "
          f"{a.systems / best * 1e3:7.2f} solves/s, {nbytes / best / 1e6:7.1f} GB/s algorithmic = {nbytes / best / 1e6 / peak:.3f} of peak",
          flush=True)
    out[pack] = res
    if trace:
        cap = 4000
        buf = np.zeros(2 * cap, np.int64)
        _lib.check(_lib.lib().dp_debug_pcg_trace(_lib.ptr(batch.ws), len(systems), buf.ctypes.data, cap))
        lab, clk = buf[0::2], buf[1::2]
        k = int((clk != 0).sum())
        lab, clk = lab[40:k], clk[40:k]
        d = np.diff(clk) / 1.9e3
        stats = {}
        for l0, l1, dd in zip(lab[:-1], lab[1:], d):
            stats.setdefault((int(l0), int(l1)), []).append(dd)
        tot = sum(np.mean(v) for v in stats.values())
        for (l0, l1), v in sorted(stats.items()):
            kind = "barrier wait" if l0 % 8 == 1 and l1 % 8 == 2 else "work"
            print(f"   {names.get(l0//8, l0//8)}.{l0%8} -> {names.get(l1//8, l1//8)}.{l1%8}: {np.mean(v):8.2f} us ({100*np.mean(v)/tot:4.1f} %)  {kind}")
same = a.packed_only or all(g.iterations == w.iterations and g.res == w.res and torch.equal(g.x_hat, w.x_hat) for g, w in zip(out[True], out[False]))
print("bitwise identical:", same)
