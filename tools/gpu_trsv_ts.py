"""Tile-stream triangular solve (trsv_ts.cuh) against the sync-free solve on 3-D factors, with distinct copies of the
factor per system (a shared factor would be served from L2). Not a test, not the bench. Select a variant with DPCG_LIB.

    python tools/gpu_trsv_ts.py [--sides 128,256] [--batches 1,4,8,16] [--variants a,b]
"""
import argparse, copy, os, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

ap = argparse.ArgumentParser()
ap.add_argument("--sides", default="128,256")
ap.add_argument("--batches", default="1,4,8,16")
ap.add_argument("--variants", default="")
ap.add_argument("--upper", type=int, default=0)
ap.add_argument("--syncfree", type=int, default=1)
ap.add_argument("--position-space", type=int, default=0)
a = ap.parse_args()

if a.variants:
    for v in a.variants.split(","):
        env = dict(os.environ)
        if v != "default":
            env["DPCG_LIB"] = str(ROOT / "deeppreconditioning_b200" / "lib" / f"libdpcg_{v}.so")
        print(f"=== {v}", flush=True)
        subprocess.run([sys.executable, __file__, "--sides", a.sides, "--batches", a.batches, "--upper", str(a.upper),
                        "--syncfree", str(a.syncfree), "--position-space", str(a.position_space)], env=env, check=False)
    sys.exit(0)

import numpy as np, torch
import bench
from deeppreconditioning_b200 import precond, synthetic
from deeppreconditioning_b200.sparse import CsrMatrix

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
ev = lambda: torch.cuda.Event(enable_timing=True)
peak = bench.peaks()[0]


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = ev(), ev()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


for side in [int(s) for s in a.sides.split(",")]:
    st, _, rhs, sizes = synthetic.make_batch("poisson3d", side, [0], device=dev)
    n = sizes[0]
    T = CsrMatrix.from_spconv(st, n, "tril")
    del st
    if a.upper:
        T = T.transpose()
    x = rhs[0, :n].to(torch.float64)
    plan = precond.analyse(T, bool(a.upper), level_stream=False)
    base = precond.level_ordered_any(T, plan)
    nbytes = 12 * T.nnz + 4 * (n + 1) + 16 * n
    widths = torch.diff(plan.level_ptr)
    print(f"--- {side}^3: n {n}, nnz {T.nnz}, levels {plan.nlevels}, widest {int(widths.max())} rows, "
          f"{nbytes / 1e6:.0f} MB per solve", flush=True)
    ref = None
    if a.syncfree and side <= 128:
        ref = precond.triangular_solve(T, plan, x, algorithm="syncfree")
        best, med = timed(lambda: precond.triangular_solve(T, plan, x, algorithm="syncfree"))
        print(f"sync-free 1 system: {best * 1e3:.0f} us ({nbytes / best / 1e6:.0f} GB/s, {nbytes / best / 1e6 / peak:.3f} of peak)", flush=True)
    for nb in [int(s) for s in a.batches.split(",")]:
        systems, copies, outs = [], [], []
        for _ in range(nb):
            c = copy.copy(base)
            c.rowptr, c.col, c.val = base.rowptr.clone(), base.col.clone(), base.val.clone()
            p = copy.copy(plan)
            p.perm = plan.perm.clone()
            systems.append((T, p, x.clone())), copies.append(c), outs.append(torch.empty_like(x))
        pos = bool(a.position_space)
        if pos:
            systems = [(m, p, b[plan.perm.long()]) for m, p, b in systems]
        got = precond.triangular_solve_batch(systems, outs, algorithm="ts", copies=copies, position_space=pos)
        torch.cuda.synchronize()
        if os.environ.get("DPCG_NO_CHECK"):
            pass
        elif ref is not None:
            want = ref[plan.perm.long()] if pos else ref
            assert all(torch.equal(g, want) for g in got), "tile-stream differs from sync-free"
        elif nb > 1:
            assert all(torch.equal(g, got[0]) for g in got)
        prepared = precond.PreparedTriangularBatch(systems, outs, "ts", copies, pos)  # solve() = kernel launches only
        best, med = timed(prepared.solve)
        prepared.check()
        gbs = nb * nbytes / best / 1e6
        print(f"tile-stream{' (position space)' if pos else ''} {nb:3d} systems: {best * 1e3:8.0f} us (median {med * 1e3:.0f}), {gbs:6.0f} GB/s = {gbs / peak:.3f} of peak, "
              f"{best * 1e3 / plan.nlevels:.2f} us per level", flush=True)
        del systems, copies, outs, got
        torch.cuda.empty_cache()
