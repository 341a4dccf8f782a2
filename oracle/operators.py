"""ORACLE — duck-typed preconditioners for the reference loop (``M`` only needs ``@``, ``cg.py:61,81``)."""

from __future__ import annotations

import numpy as np
import torch

from . import ckernels
from .sparse import to_torch_csr, transpose_csr


class Identity:
    """``_construct_vanilla`` (``test.py:70-72``) without materialising ``eye``."""

    def __matmul__(self, r):
        return r.clone()


class Jacobi:
    """``_construct_jacobi`` (``test.py:74-79``): ``diag(1 / A_ii)``."""

    def __init__(self, diagonal):
        self.inv = 1 / torch.as_tensor(np.asarray(diagonal), dtype=torch.float64)

    def __matmul__(self, r):
        return self.inv * r


class FactoredMultiply:
    """``z = L @ (L.T @ r)`` in fp64 — the factored form of ``test.py:104`` (approximate inverse)."""

    def __init__(self, l_rowptr, l_col, l_val):
        self.L = to_torch_csr(l_rowptr, l_col, l_val)
        self.Lt = to_torch_csr(*transpose_csr(l_rowptr, l_col, l_val))

    def __matmul__(self, r):
        return self.L @ (self.Lt @ r)


class FactoredSolve:
    """``z = L^-T (L^-1 r)`` — the north_star's triangular-solve apply mode (IC(0))."""

    def __init__(self, l_rowptr, l_col, l_val):
        self.l = (np.asarray(l_rowptr, np.int32), np.asarray(l_col, np.int32), np.asarray(l_val, np.float64))
        self.lt = transpose_csr(*self.l)

    def __matmul__(self, r):
        y = ckernels.sptrsv_lower(*self.l, r.numpy())
        return torch.from_numpy(ckernels.sptrsv_upper(*self.lt, y))
