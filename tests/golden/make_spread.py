"""Generate tests/golden/pcg_spread.json: the iteration counts the UNMODIFIED reference loop
(/root/reference/uibk/deep_preconditioning/cg.py::preconditioned_conjugate_gradient) produces at the BASELINE sizes
(config 2/3: 316x316, config 4: 128^3), run in every way the reference itself may legitimately be run on the same
operands:

* torch CPU thread count (``torch.set_num_threads``: the BLAS dot behind ``torch.inner`` and the CSR ``@`` split their
  work by thread, so the summation order of the reductions depends on it);
* storage of ``A``: sparse CSR (what scales) or, once, sparse COO (another summation order of ``A @ p``); the dense ``A`` that
  ``test.py:65-68`` builds needs N^2 doubles and is impossible at these sizes;
* form of ``M``: the factored ``L @ (L.T @ r)`` in fp64, or the explicit product ``fp32(L) @ fp32(L).T`` widened to fp64
  (``test.py:104-105``; formed with scipy's sparse fp32 product here because the dense fp32 matmul needs N^2 floats).

cg.py is executed as it is; only its *inputs* vary. The spread [min, max] of the counts is the reference's own
reproducibility at that size: the CUDA path is held to [min - 1, max + 1] (north_star's +-1 applied to the band), and
to exactly +-1 wherever the band is a single value (tests/test_gpu_parity.py::test_pcg_at_baseline_sizes).

Run in the build container only (the reference is not shipped to the GPU box):
    python tests/golden/make_spread.py [--only c2|c4]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import helpers  # noqa: E402
from oracle import operators, reference  # noqa: E402
from oracle import sparse as osp  # noqa: E402

OUT = Path(__file__).parent / "pcg_spread.json"
MAX_ITER = 20000


class ExplicitSparseFp32:
    """test.py:104-105 with the product formed sparsely: fp32(L) @ fp32(L).T, widened to fp64, applied by ``@``."""

    def __init__(self, l_rowptr, l_col, l_val):
        l32 = osp.to_scipy(l_rowptr, l_col, np.asarray(l_val, np.float32))
        prod = (l32 @ l32.T).tocsr()
        prod.sort_indices()
        assert prod.dtype == np.float32
        self.nnz = int(prod.nnz)
        self.M = osp.to_torch_csr(prod.indptr, prod.indices, prod.data.astype(np.float64))

    def __matmul__(self, r):
        return self.M @ r


def operand_a(p, storage):
    csr = osp.to_torch_csr(*p.A)
    return csr if storage == "csr" else csr.to_sparse_coo().coalesce()


def run_case(cg, out, kind, side, index, net, precond, variants):
    p = helpers.problem(kind, side, index, 0.5, net)
    runs = []
    for storage, form, threads in variants:
        torch.set_num_threads(threads)
        A = operand_a(p, storage)
        if precond == "identity":
            M = operators.Identity()
        elif precond == "jacobi":
            M = operators.Jacobi(osp.to_scipy(*p.A).diagonal())
        elif precond == "ic0_solve":
            M = operators.FactoredSolve(*helpers.ic0_factor(p))
        elif form == "factored":
            M = operators.FactoredMultiply(*p.L)
        else:
            M = ExplicitSparseFp32(*p.L)
        t0 = time.perf_counter()
        _, iterations, info = cg.preconditioned_conjugate_gradient(A, p.b, M, max_iter=MAX_ITER)
        runs.append(dict(a_storage=storage, m_form=form, threads=threads, iterations=int(iterations),
                         seconds=round(time.perf_counter() - t0, 2)))
        print(kind, side, index, net, precond, runs[-1], flush=True)
    counts = [r["iterations"] for r in runs]
    out["cases"].append(dict(kind=kind, side=side, index=index, net=net, precond=precond, max_iter=MAX_ITER, n=p.n,
                             nnz_a=int(len(p.A[1])), nnz_l=int(len(p.L[1])) if p.L is not None else 0,
                             iterations_min=min(counts), iterations_max=max(counts), runs=runs,
                             b_checksum=float(p.b.sum()), a_checksum=float(np.sum(p.A[2])),
                             l_checksum=float(np.sum(p.L[2])) if p.L is not None else 0.0))
    OUT.write_text(json.dumps(out, indent=1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None, choices=["c2", "c4"])
    args = ap.parse_args()
    assert reference.available(), "needs /root/reference"
    cg = reference.load_cg()
    out = {"generator": "tests/golden/make_spread.py", "reference": "uibk/deep_preconditioning/cg.py:50-90",
           "torch": torch.__version__, "numpy": np.__version__, "cases": []}
    if OUT.exists() and args.only:
        out = json.loads(OUT.read_text())
        keep = "poisson3d" if args.only == "c2" else "poisson2d"
        out["cases"] = [c for c in out["cases"] if c["kind"] == keep]
    # thread counts beyond this container's 8 cores still change how torch splits its reductions (the GPU box of round 1
    # ran the same loop with 16 and 32 threads), so they are legitimate ways to run the reference
    layouts = [("csr", 1), ("csr", 2), ("csr", 4), ("csr", 8), ("csr", 16), ("csr", 32)]
    multiply = [(s, f, t) for f in ("factored", "explicit_fp32") for s, t in layouts]
    plain = [("csr", 1), ("csr", 8), ("csr", 16)]
    if args.only in (None, "c2"):
        # BASELINE configs 2 and 3: 316x316, random-init PreconditionerNet factor, multiply mode (the benchmarked technique)
        for index in (0, 1, 2):  # (torch's COO `@` is 30x slower than CSR at this size: one run, on system 0)
            run_case(cg, out, "poisson2d", 316, index, "net", "multiply", multiply + ([("coo", "factored", 8)] if index == 0 else []))
        run_case(cg, out, "poisson2d", 316, 0, "net", "identity", [(s, "-", t) for s, t in plain])
        run_case(cg, out, "poisson2d", 316, 0, "net", "jacobi", [(s, "-", t) for s, t in plain])
        run_case(cg, out, "poisson2d", 316, 0, "net", "ic0_solve", [(s, "-", t) for s, t in plain])
    if args.only in (None, "c4"):
        # BASELINE config 4: 128^3, Jacobi, tril-pattern CNN factor (multiply), IC(0) (solve)
        run_case(cg, out, "poisson3d", 128, 0, "tril", "jacobi", [(s, "-", t) for s, t in plain[:2]])
        run_case(cg, out, "poisson3d", 128, 0, "tril", "multiply", [("csr", "factored", 1), ("csr", "factored", 8),
                                                                   ("csr", "explicit_fp32", 8)])
        run_case(cg, out, "poisson3d", 128, 0, "tril", "ic0_solve", [(s, "-", t) for s, t in plain[:2]])


if __name__ == "__main__":
    main()
