"""ctypes wrapper around ``oracle/kernels.c`` (ORACLE — test infrastructure, see ``oracle/__init__.py``)."""

from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "liboracle.so"
_lib = None


def build(force: bool = False) -> Path:
    src = _HERE / "kernels.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not _SO.exists():
            build()
        _lib = ctypes.CDLL(str(_SO))
        _lib.oracle_levels.restype = ctypes.c_int
        _lib.oracle_ic0.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def spmv_csr(rowptr, col, val, x):
    rowptr, col, val, x = _i32(rowptr), _i32(col), _f64(val), _f64(x)
    n = rowptr.shape[0] - 1
    y = np.empty(n, np.float64)
    lib().oracle_spmv_csr(ctypes.c_int(n), _p(rowptr), _p(col), _p(val), _p(x), _p(y))
    return y


def levels(rowptr, col, upper: bool = False):
    """Returns ``(level[n], perm[n], level_ptr[nlev+1])`` of a triangular CSR pattern."""
    rowptr, col = _i32(rowptr), _i32(col)
    n = rowptr.shape[0] - 1
    level = np.zeros(n, np.int32)
    nlev = lib().oracle_levels(ctypes.c_int(n), _p(rowptr), _p(col), ctypes.c_int(int(upper)), _p(level))
    perm = np.empty(n, np.int32)
    level_ptr = np.empty(nlev + 1, np.int32)
    lib().oracle_level_perm(ctypes.c_int(n), _p(level), ctypes.c_int(nlev), _p(perm), _p(level_ptr))
    return level, perm, level_ptr


def sptrsv_lower(rowptr, col, val, b):
    rowptr, col, val, b = _i32(rowptr), _i32(col), _f64(val), _f64(b)
    n = rowptr.shape[0] - 1
    y = np.empty(n, np.float64)
    lib().oracle_sptrsv_lower(ctypes.c_int(n), _p(rowptr), _p(col), _p(val), _p(b), _p(y))
    return y


def sptrsv_upper(rowptr, col, val, b):
    rowptr, col, val, b = _i32(rowptr), _i32(col), _f64(val), _f64(b)
    n = rowptr.shape[0] - 1
    z = np.empty(n, np.float64)
    lib().oracle_sptrsv_upper(ctypes.c_int(n), _p(rowptr), _p(col), _p(val), _p(b), _p(z))
    return z


def ic0(rowptr, col, val):
    """IC(0) values on the pattern of ``tril(A)`` (CSR, sorted, diagonal last). Raises on breakdown."""
    rowptr, col, val = _i32(rowptr), _i32(col), _f64(val)
    n = rowptr.shape[0] - 1
    out = np.zeros_like(val)
    status = lib().oracle_ic0(ctypes.c_int(n), _p(rowptr), _p(col), _p(val), _p(out))
    if status:
        raise FloatingPointError(f"IC(0) breakdown at row {status - 1}")
    return out


def coo_spmv_batch(indices, features, vec, transpose: bool):
    indices = _i32(indices)
    features = np.ascontiguousarray(features, np.float32).reshape(-1)
    vec = np.ascontiguousarray(vec, np.float32)
    out = np.empty_like(vec)
    lib().oracle_coo_spmv_batch(ctypes.c_long(indices.shape[0]), _p(indices), _p(features),
                                ctypes.c_int(vec.shape[0]), ctypes.c_int(vec.shape[1]), _p(vec),
                                ctypes.c_int(int(transpose)), _p(out))
    return out
