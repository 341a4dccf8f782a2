"""Timing of the standalone SpMV kernel (fp64 stream and packed copy) on the 128^3 / 256^3 operators (library via DPCG_LIB)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from deeppreconditioning_b200 import synthetic
from deeppreconditioning_b200.sparse import CsrMatrix

dev = torch.device("cuda", 0)
for side in (128, 256) if len(sys.argv) < 2 else (int(sys.argv[1]),):
    st, _, rhs, sizes = synthetic.make_batch("poisson3d", side, [0], device=dev)
    n = sizes[0]
    A = CsrMatrix.from_spconv(st, n, "symmetrise")
    del st
    x = rhs[0, :n].to(torch.float64)
    y = torch.empty_like(x)
    for packed in (False, True):
        if packed and A.packed() is None:
            print(f"{side}^3 packed: no exact packed copy (tile spans >= 65536 columns)")
            continue
        for _ in range(5):
            A.matvec(x, y, packed=packed)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); A.matvec(x, y, packed=packed); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        nbytes = (6 if packed else 12) * A.nnz + 4 * (n + 1) + 16 * n
        print(f"{side}^3 {'packed' if packed else 'fp64  '}: {1e3 * best:7.1f} us, {nbytes / best / 1e6:7.1f} GB/s = {nbytes / best / 1e6 / 6451.2:.3f} of peak")
    del A, x, y
    torch.cuda.empty_cache()
