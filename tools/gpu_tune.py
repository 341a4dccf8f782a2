"""Kernel numbers for one build of libdpcg (not a test, not the bench). Select a tuning variant with DPCG_LIB.

    python tools/gpu_tune.py [--systems 32] [--max-iter 300] [--variants a.so,b.so]

Prints: SpMV {a.side3}^3 GB/s, SpTRSV {a.side3}^3, batched PCG (capped iterations) GB/s, single-system us/iteration.
With --variants it re-runs itself once per library (the library is chosen at import time).
"""
import argparse, os, subprocess, sys, time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

ap = argparse.ArgumentParser()
ap.add_argument("--systems", type=int, default=32)
ap.add_argument("--max-iter", type=int, default=300)
ap.add_argument("--variants", default="")
ap.add_argument("--side3", type=int, default=128)
ap.add_argument("--skip", default="", help="comma list of: spmv,trsv,batch,single")
a = ap.parse_args()

if a.variants:
    for v in a.variants.split(","):
        env = dict(os.environ)
        if v != "default":
            env["DPCG_LIB"] = str(ROOT / "deeppreconditioning_b200" / "lib" / f"libdpcg_{v}.so")
        print(f"=== {v}", flush=True)
        subprocess.run([sys.executable, __file__, "--systems", str(a.systems), "--max-iter", str(a.max_iter), "--skip", a.skip, "--side3", str(a.side3)],
                       env=env, check=False)
    sys.exit(0)

import numpy as np, torch
import bench
import deeppreconditioning_b200 as dp
from deeppreconditioning_b200 import precond, synthetic
from deeppreconditioning_b200.sparse import CsrMatrix

skip = set(a.skip.split(","))
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
ev = lambda: torch.cuda.Event(enable_timing=True)


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = ev(), ev()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


if not {"spmv", "trsv"} <= skip:
    st, _, rhs, sizes = synthetic.make_batch("poisson3d", a.side3, [0], device=dev)
    n3 = sizes[0]
    A3, T3 = CsrMatrix.from_spconv(st, n3, "symmetrise"), CsrMatrix.from_spconv(st, n3, "tril")
    x = rhs[0, :n3].to(torch.float64); y = torch.empty_like(x)
    if "spmv" not in skip:
        best, med = timed(lambda: A3.matvec(x, y), 20)
        byt = 12 * A3.nnz + 4 * (n3 + 1) + 16 * n3
        print(f"spmv {a.side3}^3: best {best*1e3:.1f} us ({byt/best/1e6:.0f} GB/s), median {med*1e3:.1f} us ({byt/med/1e6:.0f} GB/s)", flush=True)
        best, med = timed(lambda: T3.matvec(x, y), 20)
        byt = 12 * T3.nnz + 4 * (n3 + 1) + 16 * n3
        print(f"spmv tril {a.side3}^3 (4/row): best {best*1e3:.1f} us ({byt/best/1e6:.0f} GB/s)", flush=True)
    if "trsv" not in skip:
        fwd3 = precond.analyse(T3, False)
        best, med = timed(lambda: precond.triangular_solve(T3, fwd3, x, y), 5)
        byt = 12 * T3.nnz + 4 * (n3 + 1) + 16 * n3
        print(f"sptrsv {a.side3}^3: best {best*1e3:.1f} us ({byt/best/1e6:.0f} GB/s), {best*1e3/fwd3.nlevels:.2f} us/level", flush=True)
        for nb in (4, 8, 16):
            sys_ = [(T3, fwd3, x)] * nb
            outs = [torch.empty_like(x) for _ in range(nb)]
            best, med = timed(lambda: precond.triangular_solve_batch(sys_, outs), 3)
            print(f"sptrsv batch {nb} x {a.side3}^3 (same operands, distinct outputs): best {best*1e3:.1f} us ({nb*byt/best/1e6:.0f} GB/s)", flush=True)
    del st, A3, T3, x, y

args = argparse.Namespace(systems_per_gpu=a.systems, side=316, net="net")
if "batch" not in skip or "single" not in skip:
    mine, host = bench.build_host_systems(args, 0, 1, dev)
if "batch" not in skip:
    bench.MAX_ITER = a.max_iter
    batch = bench.device_batch(host, dev)
    byt = sum(bench.iter_bytes(h["n"], h["a"][1].numel(), h["l"][1].numel()) for h in host)

    def go():
        batch.reset(); batch.solve()

    best, med = timed(go, 3)
    its = [r.iterations for r in batch.results()]
    print(f"pcg batch {a.systems} x 316^2 net-L multiply, {min(its)}..{max(its)} iters: {best:.1f} ms, "
          f"{byt*np.mean(its)/best/1e6:.0f} GB/s algorithmic ({byt*np.mean(its)/best/1e6/6451.2:.3f} of measured peak)", flush=True)
    del batch
if "single" not in skip:
    bench.MAX_ITER = 20000
    h = host[0]
    A = CsrMatrix.from_arrays(*h["a"], device=dev); L = CsrMatrix.from_arrays(*h["l"], device=dev); b = h["b"].to(dev)
    st, _, _, _ = synthetic.make_batch("poisson2d", 316, [h["index"]], device=dev)
    T = CsrMatrix.from_spconv(st, A.n, "tril")
    fwd = precond.analyse(T, False)
    factor = precond.incomplete_cholesky0(T, fwd)
    ic = dp.FactoredSolve(factor, None, fwd)
    ic_sf = dp.FactoredSolve(factor, None, fwd, level_stream=False)
    r0 = b.clone(); out = torch.empty_like(r0)
    fplan = precond.analyse(factor, False)
    for nm, alg in (("level-stream", "ls"), ("sync-free", "syncfree")):
        best, med = timed(lambda: precond.triangular_solve(factor, fplan, r0, out, algorithm=alg), 5)
        print(f"sptrsv 316^2 IC(0) factor, {nm}: {best*1e3:.1f} us ({best*1e3/fplan.nlevels:.3f} us/level)", flush=True)
    import copy

    def clone_system():  # distinct memory for every system of the batch (a shared factor makes all CTAs hit the same L2 lines)
        f = CsrMatrix(factor.rowptr.clone(), factor.col.clone(), factor.val.clone(), factor.n)
        p = copy.copy(fplan)
        p.perm = fplan.perm.clone()
        p.ls = copy.copy(fplan.ls)
        for name in ("rowptr", "col", "val", "level_sorted"):
            setattr(p.ls, name, getattr(fplan.ls, name).clone())
        p.ls.source = type(p.ls).key(f)
        return (f, p, r0.clone())

    pool = [clone_system() for _ in range(592)]
    for nb in (64, 128, 296, 592):
        sys_ = pool[:nb]
        outs = [torch.empty_like(r0) for _ in range(nb)]
        byt1 = 12 * factor.nnz + 4 * (A.n + 1) + 16 * A.n
        best, med = timed(lambda: precond.triangular_solve_batch(sys_, outs, algorithm="ls"), 3)
        print(f"sptrsv level-stream batch {nb} x 316^2: {best*1e3:.1f} us ({nb*byt1/best/1e6:.0f} GB/s algorithmic)", flush=True)
    for name, M in [("cnn_multiply", dp.FactoredMultiply(L)), ("jacobi", dp.Jacobi(A)), ("ic0_solve", ic), ("ic0_solve_syncfree", ic_sf)]:
        batch = dp.PcgBatch([(A, b, M)], 1e-8, 20000)

        def go1():
            batch.reset(); batch.solve()

        best, med = timed(go1, 3)
        r = batch.results()[0]
        print(f"single 316^2 {name}: {r.iterations} it, {best:.2f} ms, {best*1e3/max(r.iterations,1):.2f} us/it", flush=True)
