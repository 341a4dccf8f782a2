"""ORACLE — loader for the unmodified reference loop (build container only; never on the GPU box)."""

from __future__ import annotations

import importlib.util
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")
_CG = REFERENCE_ROOT / "uibk" / "deep_preconditioning" / "cg.py"


def available() -> bool:
    return _CG.exists()


def load_cg():
    """Import ``uibk/deep_preconditioning/cg.py`` by path (it depends on torch only, ``cg.py:6-12``)."""
    spec = importlib.util.spec_from_file_location("_reference_cg", _CG)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module
