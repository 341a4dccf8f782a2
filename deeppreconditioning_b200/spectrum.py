"""Condition-number estimate of the preconditioned operator from the CG coefficients.

The reference computes ``torch.linalg.cond(preconditioner @ matrix)`` on dense N x N operands
(``uibk/deep_preconditioning/test.py:111-113``), which is O(N^3) and impossible beyond a few thousand unknowns.
The PCG recurrence already carries the spectrum: with ``a_k`` (``cg.py:78``) and ``beta_k`` (``cg.py:82``) the
Lanczos tridiagonal of ``M A`` (in the ``M^-1`` inner product) is

    T[0,0] = 1/a_0,   T[j,j] = 1/a_j + beta_{j-1}/a_{j-1},   T[j,j-1] = T[j-1,j] = sqrt(beta_{j-1})/a_{j-1}

and its extreme eigenvalues (Ritz values) converge to ``lambda_max`` / ``lambda_min`` of ``M A`` from inside. For
symmetric positive definite ``M`` and ``A`` the ratio is the condition number of the symmetrically preconditioned
matrix ``M^1/2 A M^1/2`` - the quantity PCG's convergence bound depends on. (``cond(M @ A)`` of the reference is the
2-norm condition number of the non-symmetric product; both coincide when ``M`` commutes with ``A``, e.g. ``M = I``.)
"""

from __future__ import annotations

import numpy as np
from scipy.linalg import eigvalsh_tridiagonal


def lanczos_tridiagonal(alphas, betas):
    """Diagonal and off-diagonal of T from ``alphas[k] = a_k`` and ``betas[k]`` = the beta that built ``p_k``
    (``betas[0]`` is ignored: ``p_0 = z_0``)."""
    a = np.asarray(alphas, dtype=np.float64)
    b = np.asarray(betas, dtype=np.float64)[: len(a)]
    if len(a) == 0:
        return np.zeros(0), np.zeros(0)
    diag = 1.0 / a
    diag[1:] += b[1:] / a[:-1]
    off = np.sqrt(np.maximum(b[1:], 0.0)) / a[:-1]
    return diag, off


def ritz_extremes(alphas, betas):
    """``(lambda_min, lambda_max)`` of the Lanczos tridiagonal (NaN without a single complete iteration)."""
    diag, off = lanczos_tridiagonal(alphas, betas)
    ok = np.isfinite(diag)
    if len(diag) == 0 or not ok.all():
        m = int(np.argmin(ok)) if len(diag) and not ok.all() else len(diag)
        diag, off = diag[:m], off[: max(m - 1, 0)]
    if len(diag) == 0:
        return float("nan"), float("nan")
    if len(diag) == 1:
        return float(diag[0]), float(diag[0])
    lo = eigvalsh_tridiagonal(diag, off, select="i", select_range=(0, 0))[0]
    hi = eigvalsh_tridiagonal(diag, off, select="i", select_range=(len(diag) - 1, len(diag) - 1))[0]
    return float(lo), float(hi)


def kappa_estimate(alphas, betas) -> float:
    """``lambda_max / lambda_min`` of the preconditioned operator as seen by the solve that produced the coefficients."""
    lo, hi = ritz_extremes(alphas, betas)
    return hi / lo if lo > 0 else float("nan")
