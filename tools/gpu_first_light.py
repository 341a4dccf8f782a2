"""First-light diagnostics on a GPU box: every kernel against the oracle, verbose (not a test, not a bench)."""
import sys, time, traceback
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import numpy as np, torch
import deeppreconditioning_b200 as dp
from deeppreconditioning_b200 import _lib, precond
from deeppreconditioning_b200.sparse import CsrMatrix
from oracle import ckernels, pcg as opcg, operators as oop, sparse as osp
import helpers

dev = torch.device("cuda", 0)
print(torch.cuda.get_device_name(0), _lib.device_info(), flush=True)

def step(name, fn):
    t = time.time()
    try:
        r = fn(); torch.cuda.synchronize()
        print(f"[ok ] {name}: {r}  ({time.time()-t:.2f}s)", flush=True)
    except Exception as e:
        print(f"[ERR] {name}: {e!r}", flush=True); traceback.print_exc()

for kind, side, net in [("poisson2d", 16, "net"), ("poisson2d", 64, "net"), ("poisson3d", 12, "tril"), ("poisson2d", 100, "tril")]:
    print(f"==== {kind} {side} {net}", flush=True)
    p = helpers.problem(kind, side, 0, 0.5, net)
    st = helpers.to_device(p.systems_tril, dev); ln = helpers.to_device(p.learned, dev)
    box = {}
    def asm():
        box["A"] = CsrMatrix.from_spconv(st, p.n, "symmetrise"); helpers.assert_csr_equal(box["A"], p.A)
        box["T"] = CsrMatrix.from_spconv(st, p.n, "tril"); helpers.assert_csr_equal(box["T"], p.T)
        box["L"] = CsrMatrix.from_spconv(ln, p.n, "tril"); helpers.assert_csr_equal(box["L"], p.L)
        box["Lt"] = CsrMatrix.from_spconv(ln, p.n, "tril_t"); helpers.assert_csr_equal(box["Lt"], osp.transpose_csr(*p.L))
        helpers.assert_csr_equal(box["L"].transpose(), osp.transpose_csr(*p.L))
        return f"nnzA={box['A'].nnz} nnzL={box['L'].nnz}"
    step("assembly bit-exact", asm)
    def spmv():
        x = torch.from_numpy(np.random.default_rng(1).standard_normal(p.n)).to(dev)
        out = []
        for nm, m, o in [("A", box["A"], p.A), ("L", box["L"], p.L)]:
            y = m.matvec(x).cpu().numpy(); yo = ckernels.spmv_csr(*o, x.cpu().numpy())
            out.append((nm, bool(np.array_equal(y, yo)), float(np.abs(y-yo).max())))
        return out
    step("spmv bit-exact", spmv)
    def lev():
        out = []
        for upper, M, o in [(False, box["T"], p.T), (True, box["T"].transpose(), osp.transpose_csr(*p.T))]:
            pl = precond.analyse(M, upper); box["plan%d" % upper] = pl
            level, perm, lp = ckernels.levels(o[0], o[1], upper)
            out.append((upper, pl.nlevels, len(lp)-1, bool(np.array_equal(pl.level.cpu().numpy(), level)), bool(np.array_equal(pl.perm.cpu().numpy(), perm)), bool(np.array_equal(pl.level_ptr.cpu().numpy(), lp)), pl.nchunks, pl.max_level_chunks))
        return out
    step("levels bit-exact", lev)
    def ic0():
        Lic = precond.incomplete_cholesky0(box["T"], box["plan0"]); box["Lic"] = Lic
        want = ckernels.ic0(*p.T)
        got = Lic.val.cpu().numpy()
        return bool(np.array_equal(got, want)), float(np.abs(got-want).max())
    step("ic0 bit-exact", ic0)
    def trsv():
        Lic = box["Lic"]; b = p.b.to(dev)
        y = precond.triangular_solve(Lic, box["plan0"], b)
        yo = ckernels.sptrsv_lower(p.T[0], p.T[1], Lic.val.cpu().numpy(), p.b.numpy())
        Lt = Lic.transpose(); z = precond.triangular_solve(Lt, box["plan1"], y)
        lt = osp.transpose_csr(p.T[0], p.T[1], Lic.val.cpu().numpy())
        zo = ckernels.sptrsv_upper(*lt, yo)
        return bool(np.array_equal(y.cpu().numpy(), yo)), bool(np.array_equal(z.cpu().numpy(), zo))
    step("sptrsv bit-exact", trsv)
    At = osp.to_torch_csr(*p.A)
    for engine in ("stepped", "fused"):
        for name, Mo, Mg in [
            ("identity", oop.Identity(), lambda: dp.Identity()),
            ("jacobi", oop.Jacobi(osp.to_scipy(*p.A).diagonal()), lambda: dp.Jacobi(box["A"])),
            ("multiply", oop.FactoredMultiply(*p.L), lambda: dp.FactoredMultiply(box["L"], box["Lt"])),
            ("csr", osp.explicit_product(*p.L) if p.n <= 5000 else None, lambda: dp.CsrOperator(osp.explicit_product(*p.L), dev)),
            ("solve", oop.FactoredSolve(p.T[0], p.T[1], ckernels.ic0(*p.T)), lambda: dp.FactoredSolve(box["Lic"], None, box["plan0"], box["plan1"])),
        ]:
            if Mo is None: continue
            def run():
                o = opcg.preconditioned_conjugate_gradient(At, p.b, Mo, max_iter=3000)
                r = dp.pcg_solve(box["A"], p.b.to(dev), Mg(), max_iter=3000, engine=engine, history=True)
                xerr = float((r.x_hat.cpu() - o.x_hat).norm() / o.x_hat.norm())
                return f"iters gpu={r.iterations} oracle={o.iterations} res gpu={r.res:.3e} oracle={o.res:.3e} xrel={xerr:.2e} t={r.seconds*1e3:.2f}ms"
            step(f"pcg {engine} {name}", run)
print("done", flush=True)
