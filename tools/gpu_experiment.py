"""Knob sweeps on a resident batch (not a test, not the bench)."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import argparse, numpy as np, torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--systems", type=int, default=32)
ap.add_argument("--max-iter", type=int, default=300)
ap.add_argument("--knob", default="DPCG_CARVEOUT")
ap.add_argument("--values", default="default,25,50,75,100")
a = ap.parse_args()
bench.MAX_ITER = a.max_iter
args = argparse.Namespace(systems_per_gpu=a.systems, side=316, net="net")
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
mine, host = bench.build_host_systems(args, 0, 1, dev)
batch = bench.device_batch(host, dev)
byt = sum(bench.iter_bytes(h["n"], h["a"][1].numel(), h["l"][1].numel()) for h in host)
for v in a.values.split(","):
    if v == "default":
        os.environ.pop(a.knob, None)
    else:
        os.environ[a.knob] = v
    ts = []
    for _ in range(3):
        batch.reset(); torch.cuda.synchronize(); t0 = time.perf_counter(); batch.solve(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    its = [r.iterations for r in batch.results()]
    t = min(ts)
    print(f"{a.knob}={v}: {t*1e3:.1f} ms, iters {min(its)}..{max(its)}, {byt*np.mean(its)/t/1e9:.0f} GB/s algorithmic", flush=True)
