import sys, time, cProfile, pstats, io, argparse
sys.path.insert(0, "/root/repo")
import torch, bench
import deeppreconditioning_b200 as dp
from deeppreconditioning_b200.sparse import CsrMatrix
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
bench.MAX_ITER = 50

args = argparse.Namespace(side=316, net="net", systems_total=32, step_systems=32)
_, host = bench.build_chunk(args, list(range(32)), bench.make_net(args, dev), dev, keep_host=True)
def e2e_step():
    systems = []
    for h in host:
        A = CsrMatrix.from_arrays(*h["a"], device=dev)
        L = CsrMatrix.from_arrays(*h["l"], device=dev)
        systems.append((A, h["b"], dp.FactoredMultiply(L)))
    return dp.pcg_solve_batch(systems, 1e-8, 50, device=dev)
e2e_step(); torch.cuda.synchronize()
t0 = time.perf_counter(); e2e_step(); torch.cuda.synchronize(); print("e2e step (32 systems, 50 iterations):", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); e2e_step(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:3500])
