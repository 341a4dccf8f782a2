"""Latency probe (not a test, not the bench): single-system solves at BASELINE config-2 size with the CTA-0 timeline."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ["DPCG_TRACE"] = "1"
import numpy as np, torch
import deeppreconditioning_b200 as dp
from deeppreconditioning_b200 import _lib, model as models, precond, synthetic
from deeppreconditioning_b200.sparse import CsrMatrix

dev = torch.device("cuda", 0)
side = int(sys.argv[1]) if len(sys.argv) > 1 else 316
st, _, rhs, sizes = synthetic.make_batch("poisson2d", side, [0], device=dev)
n = sizes[0]
torch.manual_seed(69)
with torch.no_grad():
    Ln = models.PreconditionerNet(models.DEFAULT_CHANNELS).to(dev)(st)
    Lt_ = models.PreconditionerTrilNet(models.DEFAULT_CHANNELS).to(dev)(st)
A = CsrMatrix.from_spconv(st, n, "symmetrise"); T = CsrMatrix.from_spconv(st, n, "tril")
b = rhs[0, :n].to(torch.float64)
fwd = precond.analyse(T, False); ic = precond.incomplete_cholesky0(T, fwd)
names = {1: "A", 2: "APPLY1", 3: "APPLY2"}
for name, M in [("identity", dp.Identity()), ("jacobi", dp.Jacobi(A)),
                ("multiply_tril", dp.FactoredMultiply(CsrMatrix.from_spconv(Lt_, n, "tril"))),
                ("multiply_net", dp.FactoredMultiply(CsrMatrix.from_spconv(Ln, n, "tril"))),
                ("ic0_solve", dp.FactoredSolve(ic, None, fwd))]:
    batch = dp.PcgBatch([(A, b, M)], 1e-8, 20000)
    for _ in range(2):
        batch.reset(); torch.cuda.synchronize(); t0 = time.perf_counter(); batch.solve(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    r = batch.results()[0]
    print(f"{name}: {r.iterations} it, {dt*1e3:.2f} ms, {dt*1e6/max(r.iterations,1):.2f} us/it", flush=True)
    cap = 400
    out = np.zeros(2 * cap, np.int64)
    _lib.check(_lib.lib().dp_debug_pcg_trace(_lib.ptr(batch.ws), 1, out.ctypes.data, cap))
    lab, clk = out[0::2], out[1::2]
    k = int((clk != 0).sum())
    if k > 40:
        lab, clk = lab[20:k], clk[20:k]  # skip the first iterations
        d = np.diff(clk) / 1.965e3  # us at 1965 MHz
        stats = {}
        for l0, l1, dd in zip(lab[:-1], lab[1:], d):
            stats.setdefault((int(l0), int(l1)), []).append(dd)
        for (l0, l1), v in sorted(stats.items()):
            print(f"   {names.get(l0//8, l0//8)}.{l0%8} -> {names.get(l1//8, l1//8)}.{l1%8}: {np.mean(v):.2f} us (n={len(v)})")
