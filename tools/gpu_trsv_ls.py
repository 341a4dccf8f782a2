"""Level-stream triangular solve (trsv_ls.cuh) on the IC(0) factor of 2-D systems: one solve, a batch of independent
solves (distinct copies of the factor), natural numbering and level order (perm = identity: coalesced vectors), and
IC(0)-PCG on one system. Not a test, not the bench.

    python tools/gpu_trsv_ls.py [--side 316] [--batch 128]
"""
import argparse, copy, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, default=316)
ap.add_argument("--batch", type=int, default=128)
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


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = ev(), ev()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


st, _, rhs, sizes = synthetic.make_batch("poisson2d", a.side, [0], device=dev)
n = sizes[0]
b = rhs[0, :n].to(torch.float64)
for order_name in ("natural", "level"):
    st_o, b_o = st, b
    T = CsrMatrix.from_spconv(st, n, "tril")
    if order_name == "level":
        order = precond.level_ordering(T)
        st_o = order.renumber(st)
        b_o = order.to_level(b)
        T = CsrMatrix.from_spconv(st_o, n, "tril")
    A = CsrMatrix.from_spconv(st_o, n, "symmetrise")
    factor = precond.incomplete_cholesky0(T)
    plan = precond.analyse(factor, False)
    assert plan.ls is not None, "factor does not qualify for the level-stream solve"
    nbytes = 12 * factor.nnz + 4 * (n + 1) + 16 * n
    y = torch.empty_like(b_o)
    ref = precond.triangular_solve(factor, plan, b_o, algorithm="syncfree")
    got = precond.triangular_solve(factor, plan, b_o, y, algorithm="ls")
    assert torch.equal(ref, got), "level-stream differs from sync-free"
    one = precond.PreparedTriangularBatch([(factor, plan, b_o)], [y], "ls")  # solve() = the kernel launch only
    for alg in ("ls", "syncfree"):
        ms = timed(one.solve if alg == "ls" else (lambda: precond.triangular_solve(factor, plan, b_o, y, algorithm=alg)))
        print(f"[{order_name} order] {alg:8s} one {a.side}^2 solve: {1e3 * ms:7.1f} us, {1e3 * ms / plan.nlevels:.3f} us per level "
              f"({plan.nlevels} levels), {nbytes / ms / 1e6:.0f} GB/s", flush=True)

    def clone_system():
        f = CsrMatrix(factor.rowptr.clone(), factor.col.clone(), factor.val.clone(), factor.n)
        p = copy.copy(plan)
        p.perm = plan.perm.clone()
        p.ls = copy.copy(plan.ls)
        for fld in ("rowptr", "col", "val", "level_sorted"):
            setattr(p.ls, fld, getattr(plan.ls, fld).clone())
        p.ls.source = type(p.ls).key(f)
        return (f, p, b_o.clone())

    for nb in sorted({16, a.batch}):
        systems = [clone_system() for _ in range(nb)]
        outs = [torch.empty_like(b_o) for _ in range(nb)]
        got = precond.triangular_solve_batch(systems, outs, algorithm="ls")
        torch.cuda.synchronize()
        assert all(torch.equal(g, ref) for g in got)
        ms = timed(precond.PreparedTriangularBatch(systems, outs, "ls").solve, reps=3)
        gbs = nb * nbytes / ms / 1e6
        print(f"[{order_name} order] level-stream batch of {nb:3d}: {1e3 * ms:7.1f} us, {gbs:6.0f} GB/s = {gbs / peak:.3f} of peak", flush=True)
        del systems, outs, got
    for name, M in (("ic0 level-stream", dp.FactoredSolve(factor, None, plan)),
                    ("ic0 sync-free", dp.FactoredSolve(factor, None, plan, level_stream=False)), ("jacobi", dp.Jacobi(A))):
        batch = dp.PcgBatch([(A, b_o, M)], bench.RTOL, bench.MAX_ITER)

        def go():
            batch.reset()
            batch.solve()

        ms = timed(go, reps=3)
        r = batch.results()[0]
        print(f"[{order_name} order] PCG {name:18s}: {ms:7.2f} ms, {r.iterations} iterations, {1e3 * ms / r.iterations:.1f} us per iteration", flush=True)
