"""Generate tests/golden/pcg_golden.json by running the UNMODIFIED reference loop
(/root/reference/uibk/deep_preconditioning/cg.py::preconditioned_conjugate_gradient) on seeded synthetic systems.

Run in the build container only (the reference is not shipped to the GPU box):
    python tests/golden/make_golden.py
The operands are exactly what tests/helpers.py rebuilds from the same seeds, so the CUDA path and the oracle can be
compared with these numbers without the reference being present.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import helpers  # noqa: E402
from oracle import ckernels, operators, reference  # noqa: E402
from oracle import sparse as osp  # noqa: E402

CASES = [
    # kind, side, net, preconditioners, max_iter
    ("poisson2d", 16, "net", ["identity", "jacobi", "multiply", "explicit", "ic0_solve"], 3000),
    ("poisson2d", 64, "net", ["identity", "jacobi", "multiply", "explicit", "ic0_solve"], 3000),   # BASELINE config 1
    ("poisson2d", 64, "tril", ["multiply", "explicit"], 3000),
    ("poisson3d", 12, "tril", ["identity", "jacobi", "multiply", "ic0_solve"], 3000),
    ("poisson2d", 100, "tril", ["identity", "jacobi", "multiply", "ic0_solve"], 3000),
    ("poisson2d", 64, "net", ["identity"], 50),  # saturates max_iter
]


def build_operator(p, name):
    if name == "identity":
        return operators.Identity()
    if name == "jacobi":
        return operators.Jacobi(osp.to_scipy(*p.A).diagonal())
    if name == "multiply":
        return operators.FactoredMultiply(*p.L)
    if name == "explicit":
        return osp.explicit_product(*p.L)  # literal test.py:104-105
    if name == "ic0_solve":
        return operators.FactoredSolve(*helpers.ic0_factor(p))
    raise KeyError(name)


def main():
    assert reference.available(), "needs /root/reference"
    cg = reference.load_cg()
    out = {"generator": "tests/golden/make_golden.py", "reference": "uibk/deep_preconditioning/cg.py:50-90",
           "torch": torch.__version__, "numpy": np.__version__, "cases": []}
    for kind, side, net, names, max_iter in CASES:
        p = helpers.problem(kind, side, 0, 0.5, net)
        A = osp.to_torch_csr(*p.A)
        for name in names:
            M = build_operator(p, name)
            _, iterations, info = cg.preconditioned_conjugate_gradient(A, p.b, M, max_iter=max_iter)
            # the same operands through the dense A the harness really builds (test.py:65-68), small cases only
            dense_iterations = None
            if p.n <= 4096 and name in ("identity", "jacobi"):
                _, dense_iterations, _ = cg.preconditioned_conjugate_gradient(A.to_dense(), p.b, M, max_iter=max_iter)
            out["cases"].append(dict(kind=kind, side=side, net=net, precond=name, max_iter=max_iter, n=p.n,
                                     nnz_a=int(len(p.A[1])), nnz_l=int(len(p.L[1])), iterations=int(iterations),
                                     info=int(info), dense_iterations=dense_iterations,
                                     b_checksum=float(p.b.sum()), a_checksum=float(np.sum(p.A[2])),
                                     l_checksum=float(np.sum(p.L[2]))))
            print(out["cases"][-1])
    # level-set depth of tril(A) in natural ordering: 2n-1 / 3n-2
    out["levels"] = []
    for kind, side in [("poisson2d", 16), ("poisson2d", 64), ("poisson3d", 12)]:
        p = helpers.problem(kind, side, 0, 0.5, None)
        _, _, lp = ckernels.levels(p.T[0], p.T[1])
        out["levels"].append(dict(kind=kind, side=side, nlevels=int(len(lp) - 1)))
    (Path(__file__).parent / "pcg_golden.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
