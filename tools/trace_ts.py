"""Per-tile timeline of the short-row tile-stream solve (needs libdpcg_tstrace.so built with -DDPCG_TS_TRACE:
`python -c "from deeppreconditioning_b200 import build; build.build_variant('tstrace', {'DPCG_TS_TRACE': 1})"`).

    python tools/trace_ts.py [--side 256] [--batch 8]
"""
import argparse, copy, os, sys, ctypes
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ.setdefault("DPCG_LIB", str(ROOT / "deeppreconditioning_b200" / "lib" / "libdpcg_tstrace.so"))
import numpy as np, torch
from deeppreconditioning_b200 import precond, synthetic, _lib
from deeppreconditioning_b200.sparse import CsrMatrix

ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, default=256)
ap.add_argument("--batch", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda", 0)
st, _, rhs, sizes = synthetic.make_batch("poisson3d", a.side, [0], device=dev)
n = sizes[0]
T = CsrMatrix.from_spconv(st, n, "tril")
del st
x = rhs[0, :n].to(torch.float64)
plan = precond.analyse(T, False, level_stream=False)
base = precond.level_ordered_any(T, plan)
systems, copies, outs = [], [], []
for _ in range(a.batch):
    c = copy.copy(base)
    c.rowptr, c.col, c.val = base.rowptr.clone(), base.col.clone(), base.val.clone()
    p = copy.copy(plan)
    systems.append((T, p, x[plan.perm.long()])), copies.append(c), outs.append(torch.empty_like(x))
for _ in range(3):
    precond.triangular_solve_batch(systems, outs, algorithm="ts", copies=copies, position_space=True)
torch.cuda.synchronize()
K = 1024
out = np.zeros(2 * K * 8, np.int64)
h = _lib.lib(); h.dp_debug_ts_trace.argtypes = [ctypes.c_void_p]; h.dp_debug_ts_trace(out.ctypes.data)
t = out.reshape(2, K, 8)
names = ["wait for the stage's bytes", "stage -> registers", "release", "loads issued, next tile's bytes awaited, first try", "wait for a level"]
for cta in (0, 1):
    used = int((t[cta, :, 0] > 0).sum())
    if used < 8:
        continue
    tt = t[cta, :used, :6]
    d = np.diff(tt, axis=1)
    mid = slice(used // 4, 3 * used // 4)
    print(f"CTA {'0' if cta == 0 else 'middle'}: {used} tiles traced; tile period mean {np.diff(tt[mid, 0]).mean():.0f} cycles "
          f"(first quarter {np.diff(tt[:used // 4, 0]).mean():.0f}, last quarter {np.diff(tt[3 * used // 4:, 0]).mean():.0f})")
    print("   mean cycles (middle half):", {names[i]: round(float(d[mid, i].mean())) for i in range(5)})
    gap = tt[1:, 0] - tt[:-1, 5]
    print("   tile end -> next tile start:", round(float(gap[mid].mean())))
    for i in range(used // 2, used // 2 + 6):
        print("   tile", i, d[i].tolist())
