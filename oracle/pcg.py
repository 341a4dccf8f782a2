"""ORACLE — restatement of ``uibk/deep_preconditioning/cg.py`` (test infrastructure, see ``oracle/__init__.py``).

Follows the reference line by line and with the same torch CPU operators (``@``, ``torch.inner``) so the
floating-point history is the reference's; the only addition is that it hands back what the reference
computes but drops (``x_hat``, the last ``res``, the residual history, the coefficients ``a``/``beta``) —
``cg.py:90`` returns only ``(seconds, iterations, 0)``.
"""

from __future__ import annotations

import time
from dataclasses import dataclass, field

import torch


def stopping_criterion(_, rk, b):
    """cg.py:15-17 — the *squared* relative residual."""
    return torch.inner(rk, rk) / torch.inner(b, b)


@dataclass
class PcgResult:
    seconds: float
    iterations: int
    info: int
    x_hat: torch.Tensor
    res: float
    history: list = field(default_factory=list)
    alphas: list = field(default_factory=list)  # a of every body (cg.py:78)
    betas: list = field(default_factory=list)   # beta behind every body's p (cg.py:82 of the previous body; 0 first)


def preconditioned_conjugate_gradient(A, b, M, x0=None, x_true=None, rtol=1e-8, max_iter=1024, as_is=False) -> PcgResult:
    """cg.py:50-90. ``M`` is applied by ``@`` (an approximate inverse).

    ``as_is``: also execute the work the reference does and drops every iteration - the error vector (all zeros when
    ``x_true`` is None) and its A-norm ``<e, A e>``, a second ``A @`` per body (cg.py:85,87). It never influences the
    iteration; the timed CPU baseline of ``bench.py`` switches it on so that the loop is timed as the reference runs it."""
    x_hat = x0 if x0 is not None else torch.zeros_like(b, dtype=torch.float64)  # cg.py:58

    rk = b - A @ x_hat  # cg.py:60
    zk = M @ rk  # cg.py:61
    pk = zk.clone()  # cg.py:62

    res = stopping_criterion(A, zk, b)  # cg.py:66 — iteration 0 checks the PRECONDITIONED residual
    history = [res.item()]
    alphas, betas, beta = [], [], torch.zeros((), dtype=torch.float64)

    start_time = time.perf_counter()  # cg.py:69
    for _ in range(max_iter):  # cg.py:70
        if res < rtol:  # cg.py:71
            break
        Ap = A @ pk  # cg.py:75
        rz = torch.inner(rk, zk)  # cg.py:76
        a = rz / torch.inner(Ap, pk)  # cg.py:78
        alphas.append(a.item()), betas.append(beta.item())
        x_hat = x_hat + a * pk  # cg.py:79
        rk = rk - a * Ap  # cg.py:80
        zk = M @ rk  # cg.py:81
        beta = torch.inner(rk, zk) / rz  # cg.py:82
        pk = zk + beta * pk  # cg.py:83
        if as_is:
            error_i = (x_hat - x_true) if x_true is not None else torch.zeros_like(b, requires_grad=False)  # cg.py:85
        res = stopping_criterion(A, rk, b)  # cg.py:86
        if as_is:
            torch.inner(error_i, A @ error_i)  # cg.py:87: the A-norm error the reference appends and never reads
        history.append(res.item())  # cg.py:87
    end_time = time.perf_counter()  # cg.py:88

    return PcgResult(end_time - start_time, len(history) - 1, 0, x_hat, float(history[-1]), history, alphas, betas)  # cg.py:90


def conjugate_gradient(A, b, x0=None, x_true=None, rtol=1e-8, max_iter=1024):
    """cg.py:20-47 — unpreconditioned CG, returns ``(errors, x_hat)`` with ``errors = [(A-norm error, res)]``."""
    x_hat = x0 if x0 is not None else torch.zeros_like(b)
    r = b - A @ x_hat
    p = r.clone()
    error_i = (x_hat - x_true) if x_true is not None else torch.zeros_like(b)
    res = stopping_criterion(A, r, b)
    errors = [(torch.inner(error_i, A @ error_i), res)]
    for _ in range(max_iter):
        if res < rtol:
            break
        Ap = A @ p
        r_norm = torch.inner(r, r)
        a = r_norm / torch.inner(Ap, p)
        x_hat = x_hat + a * p
        r = r - a * Ap
        p = r + (torch.inner(r, r) / r_norm) * p
        error_i = (x_hat - x_true) if x_true is not None else torch.zeros_like(b)
        res = stopping_criterion(A, r, b)
        errors.append((torch.inner(error_i, A @ error_i), res))
    return errors, x_hat
