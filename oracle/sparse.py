"""ORACLE — CPU restatement of the operand construction around the loop (test infrastructure).

* :func:`tril_coo_to_csr`     ``preconditioners_tril.dense()[0,0,:n,:n]`` → ``to_sparse_csr()`` without the product
                              (``test.py:102-105``): keep ``row >= col`` and ``value != 0``, fp32 → fp64 exactly.
* :func:`symmetrise_tril`     ``matrix += torch.tril(matrix, -1).T`` (``test.py:65-68``).
* :func:`explicit_product`    the literal ``test.py:104-105``: fp32 ``L @ L.T`` widened to fp64, exact zeros dropped.
* :func:`sparse_matvec_mul`   ``utils.py:15-43``.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch

from . import ckernels


def _coo_of_batch(indices, features, batch: int, n: int):
    ind = np.asarray(indices, dtype=np.int64)
    val = np.asarray(features, dtype=np.float32).reshape(-1)
    keep = (ind[:, 0] == batch) & (ind[:, 1] < n) & (ind[:, 2] < n)
    return ind[keep, 1], ind[keep, 2], val[keep]


def _to_sorted_csr(rows, cols, vals64, n: int):
    order = np.lexsort((cols, rows))
    rows, cols, vals64 = rows[order], cols[order], vals64[order]
    if len(rows) > 1:
        assert not np.any((rows[1:] == rows[:-1]) & (cols[1:] == cols[:-1])), "duplicate (row, col)"
    rowptr = np.zeros(n + 1, np.int64)
    np.add.at(rowptr, rows + 1, 1)
    return np.cumsum(rowptr).astype(np.int32), cols.astype(np.int32), vals64.astype(np.float64)


def tril_coo_to_csr(indices, features, batch: int, n: int):
    """CSR ``(rowptr int32, col int32, val float64)`` of the lower-triangular factor, sorted by (row, col)."""
    r, c, v = _coo_of_batch(indices, features, batch, n)
    keep = (r >= c) & (v != 0)
    return _to_sorted_csr(r[keep], c[keep], v[keep].astype(np.float64), n)


def symmetrise_tril(indices, features, batch: int, n: int):
    """CSR of ``A = T + tril(T, -1)^T`` from the stored lower triangle ``T`` (zeros dropped, like ``to_sparse_csr``)."""
    r, c, v = _coo_of_batch(indices, features, batch, n)
    keep = (r >= c) & (v != 0)
    r, c, v = r[keep], c[keep], v[keep].astype(np.float64)
    strict = r > c
    return _to_sorted_csr(np.concatenate([r, c[strict]]), np.concatenate([c, r[strict]]),
                          np.concatenate([v, v[strict]]), n)


def transpose_csr(rowptr, col, val):
    n = len(rowptr) - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr))
    return _to_sorted_csr(np.asarray(col, np.int64), rows, np.asarray(val, np.float64), n)


def to_scipy(rowptr, col, val):
    n = len(rowptr) - 1
    return sp.csr_matrix((np.asarray(val), np.asarray(col), np.asarray(rowptr)), shape=(n, n))


def to_torch_csr(rowptr, col, val):
    n = len(rowptr) - 1
    return torch.sparse_csr_tensor(torch.from_numpy(np.asarray(rowptr, np.int64)),
                                   torch.from_numpy(np.asarray(col, np.int64)),
                                   torch.from_numpy(np.asarray(val, np.float64)), size=(n, n))


def explicit_product(l_rowptr, l_col, l_val):
    """``test.py:104-105``: ``L @ L.T`` in **fp32**, widened to fp64, as CSR with exact zeros dropped."""
    l32 = to_scipy(l_rowptr, l_col, np.asarray(l_val, np.float32))
    dense = torch.from_numpy(l32.toarray())
    product = torch.matmul(dense, dense.transpose(-1, -2))
    return product.to(torch.float64).to_sparse_csr()


def sparse_matvec_mul(indices, features, vector_batch, transpose: bool):
    """``utils.py:15-43`` on plain arrays: ``[B, N]`` fp32 result."""
    return ckernels.coo_spmv_batch(np.asarray(indices), np.asarray(features), np.asarray(vector_batch), transpose)


def pack_csr(rowptr, col, val, tile_rows: int = 512):
    """Restatement of ``dp_csr_pack`` (no reference counterpart: a lossless storage format of this implementation):
    per tile of ``tile_rows`` rows the smallest column and the number of columns spanned, 16-bit offsets, fp32 values. Returns
    ``(col16, val32, tile_base, status)`` with ``status`` bit 0 = a value is not an fp32 number, bit 1 = a tile spans
    65536 columns or more."""
    rowptr, col, val = np.asarray(rowptr), np.asarray(col), np.asarray(val, dtype=np.float64)
    n = len(rowptr) - 1
    ntiles = (n + tile_rows - 1) // tile_rows
    col16 = np.zeros(len(col), np.uint16)
    with np.errstate(over="ignore", invalid="ignore"):
        val32 = val.astype(np.float32)
    tile_base = np.zeros((max(ntiles, 1), 2), np.int32)
    status = 0 if np.array_equal(val32.astype(np.float64).view(np.int64), val.view(np.int64)) else 1
    for t in range(ntiles):
        cs, ce = rowptr[min(t * tile_rows, n)], rowptr[min((t + 1) * tile_rows, n)]
        if ce > cs:
            lo, hi = int(col[cs:ce].min()), int(col[cs:ce].max())
            tile_base[t] = (lo, hi - lo + 1)
            if hi - lo > 65535:
                status |= 2
            col16[cs:ce] = ((col[cs:ce] - lo) & 0xFFFF).astype(np.uint16)
    return col16, val32, tile_base, status
