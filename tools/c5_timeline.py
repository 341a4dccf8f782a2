"""Where does a config-5 iteration go? One batch solve (8 x side^3, IC(0) solve mode, level order, tile-stream solves, stepped
engine) under torch.profiler (CUPTI kernel records of the real, non-serialised run): kernel time by name vs the span of the
solve. Not a test, not the bench.   python tools/c5_timeline.py [--side 256] [--batch 8] [--max-iter 40]"""
import argparse, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, default=256)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--max-iter", type=int, default=40)
a = ap.parse_args()
import torch
from torch.profiler import ProfilerActivity, profile
import deeppreconditioning_b200 as dp
from deeppreconditioning_b200 import model as models, precond, synthetic
from deeppreconditioning_b200.sparse import CsrMatrix

dev = torch.device("cuda", 0)
systems = []
for index in range(a.batch):
    st, _, rhs, sizes = synthetic.make_batch("poisson3d", a.side, [index])
    n = sizes[0]
    st = models.SparseConvTensor(st.features.to(dev), st.indices.to(dev), st.spatial_shape, 1)
    T = CsrMatrix.from_spconv(st, n, "tril")
    order = precond.level_ordering(T)
    st = order.renumber(st)
    T = CsrMatrix.from_spconv(st, n, "tril")
    A = CsrMatrix.from_spconv(st, n, "symmetrise")
    plan = precond.analyse(T, False, level_stream=False)
    F = precond.incomplete_cholesky0(T, plan)
    b = order.to_level(rhs[0, :n].to(device=dev, dtype=torch.float64))
    systems.append((A, b, dp.FactoredSolve(F, None, plan, level_stream=False, tile_stream=True)))
batch = dp.PcgBatch(systems, 1e-8, a.max_iter, device=dev)
batch.reset(); batch.solve(); torch.cuda.synchronize()
batch.reset()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    e0.record(); batch.solve(); e1.record(); torch.cuda.synchronize()
span_ms = e0.elapsed_time(e1)
tot = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA and ev.name and "Memcpy" not in ev.name and "Memset" not in ev.name:
        k = ev.name[:60]
        c = tot.setdefault(k, [0, 0.0])
        c[0] += 1; c[1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
busy = sum(v[1] for v in tot.values()) / 1e3
print(f"{a.batch} x {a.side}^3, {a.max_iter} iterations: span {span_ms:.1f} ms ({span_ms / a.max_iter:.2f} ms per iteration), kernels {busy:.1f} ms = {100 * busy / span_ms:.0f} % of the span")
for k, (cnt, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:8]:
    print(f"  {cnt:5d} x {us / cnt:9.1f} us = {us / 1e3:8.1f} ms  {k}")
