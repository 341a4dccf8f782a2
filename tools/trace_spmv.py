"""Per-tile timeline of the standalone SpMV kernel (a -DDPCG_PIPE_TRACE build selected with DPCG_LIB), 128^3 operator."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from deeppreconditioning_b200 import _lib, synthetic
from deeppreconditioning_b200.sparse import CsrMatrix

dev = torch.device("cuda", 0)
side = int(sys.argv[1]) if len(sys.argv) > 1 else 128
st, _, rhs, sizes = synthetic.make_batch("poisson3d", side, [0], device=dev)
n = sizes[0]
A = CsrMatrix.from_spconv(st, n, "symmetrise")
x = rhs[0, :n].to(torch.float64)
y = torch.empty_like(x)
names = {(0, 6): "issue duty (thread 0)", (6, 2): "wait bytes", (2, 3): "row loop", (3, 4): "release", (4, 0): "store y, next tile head"}
for packed in (False, True):
    for _ in range(5):
        A.matvec(x, y, packed=packed)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); A.matvec(x, y, packed=packed); e1.record(); torch.cuda.synchronize()
    cap = 4096
    buf = np.zeros(2 * cap, np.uint64)
    _lib.check(_lib.lib().dp_debug_pipe_trace(buf.ctypes.data, cap, 1), "dp_debug_pipe_trace")
    print(f"{'packed' if packed else 'fp64 stream'}: {1e3 * e0.elapsed_time(e1):.1f} us per product")
    for w, who in ((0, "warp 0"), (1, "warp 5")):
        t = buf[w * cap:(w + 1) * cap]
        t = t[t != 0]
        lab, clk = (t >> np.uint64(48)).astype(np.int64) % 8, (t & np.uint64(0xFFFFFFFFFFFF)).astype(np.int64)
        first = int(np.argmax(lab == 0))
        print(f"  {who}: " + " ".join(f"{l}:{dt}" for l, dt in zip(lab[first + 1:first + 90], np.diff(clk[first:first + 90]))))
