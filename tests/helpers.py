"""Shared problem builders for the tests: the same seeded inputs feed the oracle (CPU) and the CUDA path."""

from __future__ import annotations

from dataclasses import dataclass
from functools import lru_cache

import numpy as np
import torch

from deeppreconditioning_b200 import model as models
from deeppreconditioning_b200 import synthetic
from oracle import ckernels
from oracle import sparse as osp


@dataclass
class Problem:
    kind: str
    n_side: int
    n: int
    systems_tril: models.SparseConvTensor  # tril(A), reference layout
    rhs32: torch.Tensor                    # [1, n] fp32
    b: torch.Tensor                        # [n] fp64 (fp32-rounded values)
    A: tuple                               # oracle CSR (rowptr, col, val)
    T: tuple                               # oracle CSR of tril(A)
    learned: models.SparseConvTensor | None = None  # CNN output (CPU)
    L: tuple | None = None                 # oracle CSR of the learned factor


@lru_cache(maxsize=None)
def problem(kind: str, n_side: int, index: int = 0, sigma: float = 0.5, net: str | None = None) -> Problem:
    st, _, rhs, sizes = synthetic.make_batch(kind, n_side, [index], sigma)
    n = sizes[0]
    ind, feat = st.indices.numpy(), st.features.numpy()
    p = Problem(kind, n_side, n, st, rhs, rhs[0, :n].to(torch.float64), osp.symmetrise_tril(ind, feat, 0, n),
                osp.tril_coo_to_csr(ind, feat, 0, n))
    if net is not None:
        torch.manual_seed(69)  # test.py:205
        cls = {"net": models.PreconditionerNet, "tril": models.PreconditionerTrilNet}[net]
        with torch.no_grad():
            p.learned = cls(models.DEFAULT_CHANNELS)(st)
        p.L = osp.tril_coo_to_csr(p.learned.indices.numpy(), p.learned.features.numpy(), 0, n)
    return p


def ic0_factor(p: Problem):
    return p.T[0], p.T[1], ckernels.ic0(*p.T)


def to_device(st: models.SparseConvTensor, device) -> models.SparseConvTensor:
    return models.SparseConvTensor(st.features.to(device), st.indices.to(device), st.spatial_shape, st.batch_size)


def assert_csr_equal(got, want):
    """Bit-exact comparison of a device CsrMatrix with an oracle (rowptr, col, val) triple."""
    rowptr, col, val = got.to_host()
    assert np.array_equal(rowptr, want[0]), "rowptr differs"
    assert np.array_equal(col, want[1]), "col differs"
    assert np.array_equal(val.view(np.int64), np.asarray(want[2]).view(np.int64)), "values differ (bitwise)"
