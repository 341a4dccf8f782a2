"""Generate tests/golden/metrics_golden.json with the UNMODIFIED reference losses
(/root/reference/uibk/deep_preconditioning/metrics.py:13-77) on the seeded batch of tests/test_oracle.py::_loss_batch.
Run in the build container only:  python tests/golden/make_metrics_golden.py
"""
import importlib
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, "/root/reference")

from test_oracle import _loss_batch  # noqa: E402

ref = importlib.import_module("uibk.deep_preconditioning.metrics")
st, learned, solution, rhs = _loss_batch()
torch.manual_seed(7)
out = {"generator": "tests/golden/make_metrics_golden.py", "reference": "uibk/deep_preconditioning/metrics.py:13-77",
       "frobenius_loss": float(ref.frobenius_loss(learned, solution, rhs)),
       "inverse_loss": float(ref.inverse_loss(st, learned)),
       "hutchinson_trace_seed7": float(ref.hutchinson_trace(st, learned))}
(Path(__file__).parent / "metrics_golden.json").write_text(json.dumps(out, indent=1))
print(out)
