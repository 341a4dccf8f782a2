"""Per-tile timeline of the level-stream solve (needs libdpcg_lstrace<W>.so built with -DDPCG_LS_TRACE
-DDPCG_LS_TRACE_WARP=<W>): python tools/trace_ls.py [warp ...]"""
import os, sys, ctypes, subprocess
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
if len(sys.argv) > 2 or (len(sys.argv) == 2 and "," in sys.argv[1]):
    for w in (sys.argv[1].split(",") if len(sys.argv) == 2 else sys.argv[1:]):
        subprocess.run([sys.executable, __file__, w])
    sys.exit(0)
warp = sys.argv[1] if len(sys.argv) > 1 else "0"
os.environ["DPCG_LIB"] = str(ROOT / "deeppreconditioning_b200" / "lib" / f"libdpcg_lstrace{warp}.so")
import numpy as np, torch
from deeppreconditioning_b200 import precond, synthetic, _lib
from deeppreconditioning_b200.sparse import CsrMatrix
dev = torch.device("cuda", 0)
st, _, rhs, sizes = synthetic.make_batch("poisson2d", 316, [0], device=dev)
n = sizes[0]
T = CsrMatrix.from_spconv(st, n, "tril")
order = precond.level_ordering(T)
st = order.renumber(st)
T = CsrMatrix.from_spconv(st, n, "tril")
factor = precond.incomplete_cholesky0(T)
plan = precond.analyse(factor, False)
b = order.to_level(rhs[0, :n].to(torch.float64)); x = torch.empty_like(b)
for _ in range(3):
    precond.triangular_solve(factor, plan, b, x, algorithm="ls")
torch.cuda.synchronize()
out = np.zeros(8 * 256, np.int64)
h = _lib.lib(); h.dp_debug_ls_trace.argtypes = [ctypes.c_void_p]; h.dp_debug_ls_trace(out.ctypes.data)
t = out.reshape(256, 8)
names = ["wait for the stage's bytes", "row -> registers, release", "gate (one word)", "own slots, solve, store"]
k = 190
d = np.diff(t[:k, :5], axis=1)
print(f"warp {warp}: per tile mean cycles:", {names[i]: round(float(d[20:, i].mean())) for i in range(4)})
print("   tile end -> next tile start:", round(float((t[1:k, 0] - t[:k - 1, 4])[20:].mean())), " tile period:", round(float(np.diff(t[20:k, 0]).mean())))
for i in range(100, 104):
    print("   tile", i, d[i].tolist())
