"""Per-tile timeline of the level-stream solve (needs libdpcg_lstrace.so built with -DDPCG_LS_TRACE)."""
import os, sys, ctypes
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ["DPCG_LIB"] = str(ROOT / "deeppreconditioning_b200" / "lib" / "libdpcg_lstrace.so")
import numpy as np, torch
from deeppreconditioning_b200 import precond, synthetic, _lib
from deeppreconditioning_b200.sparse import CsrMatrix
dev = torch.device("cuda", 0)
st, _, rhs, sizes = synthetic.make_batch("poisson2d", 316, [0], device=dev)
n = sizes[0]
T = CsrMatrix.from_spconv(st, n, "tril")
factor = precond.incomplete_cholesky0(T)
plan = precond.analyse(factor, False)
b = rhs[0, :n].to(torch.float64); x = torch.empty_like(b)
for _ in range(3):
    precond.triangular_solve(factor, plan, b, x, algorithm="ls")
torch.cuda.synchronize()
out = np.zeros(8 * 256, np.int64)
h = _lib.lib(); h.dp_debug_ls_trace.argtypes = [ctypes.c_void_p]; h.dp_debug_ls_trace(out.ctypes.data)
t = out.reshape(256, 8)
names = ["rhs + next tile's metadata loads", "issue_tile", "wait_item", "row -> registers, 1/diag", "level loop", "store+release"]
k = 190
d = np.diff(t[:k, :7], axis=1)
print("per tile mean cycles:", {names[i]: float(d[20:, i].mean()) for i in range(6)})
gap = t[1:k, 0] - t[:k - 1, 6]
print("tile end -> next tile start:", float(gap[20:].mean()))
print("levels per tile mean:", float(t[20:k, 7].mean()), "loop cycles per level:", float((d[20:, 4] / np.maximum(t[20:k, 7], 1)).mean()))
print("tile period:", float(np.diff(t[20:k, 0]).mean()))
for i in range(40, 48):
    print(i, d[i].tolist(), "levels", t[i, 7])
