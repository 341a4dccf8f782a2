"""Profiling driver for ncu: standalone SpMV / SpTRSV on the 128^3 system (BASELINE config 4). Not a bench."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from deeppreconditioning_b200 import precond, synthetic
from deeppreconditioning_b200.sparse import CsrMatrix

dev = torch.device("cuda", 0)
side = int(sys.argv[1]) if len(sys.argv) > 1 else 128
st, _, rhs, sizes = synthetic.make_batch("poisson3d", side, [0], device=dev)
n = sizes[0]
A, T = CsrMatrix.from_spconv(st, n, "symmetrise"), CsrMatrix.from_spconv(st, n, "tril")
x = rhs[0, :n].to(torch.float64)
y = torch.empty_like(x)
fwd = precond.analyse(T, False)
for _ in range(4):
    A.matvec(x, y)
for _ in range(3):
    precond.triangular_solve(T, fwd, x, y)
torch.cuda.synchronize()
print("done", n, A.nnz, T.nnz, fwd.nlevels)
