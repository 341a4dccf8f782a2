"""Profiling driver for ncu: level-stream triangular solve on the 316^2 IC(0) factor, one system or a batch of distinct
copies (level-ordered system: vectors by position). Not a bench.   python tools/profile_ls.py [batch]"""
import copy, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from deeppreconditioning_b200 import precond, synthetic
from deeppreconditioning_b200.sparse import CsrMatrix

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
st, _, rhs, sizes = synthetic.make_batch("poisson2d", 316, [0], device=dev)
n = sizes[0]
T = CsrMatrix.from_spconv(st, n, "tril")
order = precond.level_ordering(T)
st = order.renumber(st)
T = CsrMatrix.from_spconv(st, n, "tril")
factor = precond.incomplete_cholesky0(T)
plan = precond.analyse(factor, False)
b = order.to_level(rhs[0, :n].to(torch.float64))


def clone_system():
    f = CsrMatrix(factor.rowptr.clone(), factor.col.clone(), factor.val.clone(), factor.n)
    p = copy.copy(plan)
    p.perm = plan.perm.clone()
    p.ls = copy.copy(plan.ls)
    for fld in ("rowptr", "col", "val", "level_sorted"):
        setattr(p.ls, fld, getattr(plan.ls, fld).clone())
    p.ls.source = type(p.ls).key(f)
    return (f, p, b.clone())


systems = [clone_system() for _ in range(nb)]
outs = [torch.empty_like(b) for _ in range(nb)]
for _ in range(3):
    precond.triangular_solve_batch(systems, outs, algorithm="ls")
torch.cuda.synchronize()
print("done", n, plan.nlevels, nb)
