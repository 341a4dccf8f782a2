"""Profiling driver for ncu: level-stream triangular solve on one 316^2 IC(0) factor. Not a bench."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from deeppreconditioning_b200 import precond, synthetic
from deeppreconditioning_b200.sparse import CsrMatrix

dev = torch.device("cuda", 0)
st, _, rhs, sizes = synthetic.make_batch("poisson2d", 316, [0], device=dev)
n = sizes[0]
T = CsrMatrix.from_spconv(st, n, "tril")
factor = precond.incomplete_cholesky0(T)
plan = precond.analyse(factor, False)
b = rhs[0, :n].to(torch.float64)
x = torch.empty_like(b)
for _ in range(3):
    precond.triangular_solve(factor, plan, b, x, algorithm="ls")
torch.cuda.synchronize()
print("done", n, plan.nlevels, plan.ls is not None)
