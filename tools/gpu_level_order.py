"""IC(0)-PCG (solve mode) on 3-D systems in natural order vs level order (precond.LevelOrdering), single and batched.
Not a test, not the bench.

    python tools/gpu_level_order.py [--side 128] [--batch 8]
"""
import argparse, sys, time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, default=128)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--kind", default="poisson3d")
ap.add_argument("--modes", default="natural,level,level_ts")
ap.add_argument("--max-iter", type=int, default=0, help="cap the iterations (profiling runs)")
a = ap.parse_args()

import numpy as np, torch
import bench
import deeppreconditioning_b200 as dp
from deeppreconditioning_b200 import precond, synthetic
from deeppreconditioning_b200.sparse import CsrMatrix

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
peak = bench.peaks()[0]
ev = lambda: torch.cuda.Event(enable_timing=True)


def build(index, level, tile_stream=False):
    st, _, rhs, sizes = synthetic.make_batch(a.kind, a.side, [index], device=dev)
    n = sizes[0]
    b = rhs[0, :n].to(torch.float64)
    T = CsrMatrix.from_spconv(st, n, "tril")
    order = None
    if level:
        order = precond.level_ordering(T)
        st = order.renumber(st)
        T = CsrMatrix.from_spconv(st, n, "tril")
        b = order.to_level(b)
    A = CsrMatrix.from_spconv(st, n, "symmetrise")
    plan = precond.analyse(T, False, level_stream=False)
    F = precond.incomplete_cholesky0(T, plan)
    return A, b, dp.FactoredSolve(F, None, plan, level_stream=False, tile_stream=tile_stream), order


MODES = {"natural": (False, False), "level": (True, False), "level_ts": (True, True)}
for level, tile_stream in [MODES[m] for m in a.modes.split(",")]:
    systems, orders = [], []
    torch.cuda.empty_cache()
    for i in range(a.batch):
        A, b, M, order = build(i, level, tile_stream)
        systems.append((A, b, M)), orders.append(order)
    n, nnz_a, nnz_l = systems[0][0].n, systems[0][0].nnz, systems[0][2].L.nnz
    for nb in sorted({1, a.batch}):
        batch = dp.PcgBatch(systems[:nb], bench.RTOL, a.max_iter or bench.MAX_ITER)
        best = float("inf")
        for _ in range(2):
            batch.reset()
            e0, e1 = ev(), ev()
            e0.record(); batch.solve(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res = batch.results()
        its = sum(r.iterations for r in res)
        gbs = bench.iter_bytes(n, nnz_a, nnz_l) * its / best / 1e6
        print(f"{'level' if level else 'natural'} order{' + tile-stream solves (stepped engine)' if tile_stream else ''}, {nb} x {a.side}^3 IC(0) solve-mode PCG: {best:.1f} ms, iterations "
              f"{[r.iterations for r in res]}, {1e3 * best / max(r.iterations for r in res):.0f} us per iteration, "
              f"{gbs:.0f} GB/s = {gbs / peak:.3f} of peak", flush=True)
        del batch
