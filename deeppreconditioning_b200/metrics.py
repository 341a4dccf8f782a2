"""Sparse forms of the reference's training losses (``uibk/deep_preconditioning/metrics.py``), SURVEY §8f-3.

The reference evaluates them on dense ``n x n`` tensors (``metrics.py:44-50,66-75``), which is what confines its pipeline
to n ~ 2.4 k. Here every product is the batched COO SpMV kernel (``utils.sparse_matvec_mul`` ->
``dp_coo_spmv_batch_f32``), differentiable with respect to the CNN output, so the same quantities are available at the
BASELINE sizes:

* :func:`frobenius_loss`   ``metrics.py:13-31`` - literally two ``sparse_matvec_mul`` calls and a norm; same value.
* :func:`hutchinson_trace` ``metrics.py:58-77`` - ``mean_b || L L^T v - A v ||_2`` for a random probe ``v`` per system; same
  value for the same probes (pass ``vector=``), ``A v`` formed from the stored lower triangle as ``T v + T^T v - diag(T) v``.
* :func:`inverse_loss`     ``metrics.py:34-55`` - ``mean_b || L L^T A - I ||_F``; the Frobenius norm is taken through
  probes, ``||M||_F^2 = sum_k ||M e_k||^2``: with ``probes=None`` all ``n`` unit vectors (exact, the reference's value; small
  systems), else that many Rademacher probes (an unbiased estimate of ``||M||_F^2``, Hutchinson).
``condition_loss`` (``metrics.py:80-100``, dense SVD) is out of scope; the solve-side estimate is ``spectrum.kappa_estimate``.
"""

from __future__ import annotations

import torch

from .utils import sparse_matvec_mul


def frobenius_loss(lower_triangular, solution: torch.Tensor, right_hand_side: torch.Tensor) -> torch.Tensor:
    """Frobenius norm of the error, arXiv 2305.16432 eq. (11) (``metrics.py:13-31``): ``sum_b || L L^T x_b - b_b ||_2``."""
    interim = sparse_matvec_mul(lower_triangular, solution, transpose=True)
    interim = sparse_matvec_mul(lower_triangular, interim, transpose=False)
    return torch.linalg.vector_norm(interim - right_hand_side.to(interim.device), ord=2, dim=1).sum()


def _diagonal_part(tensor):
    """The same tensor restricted to its diagonal sites (for ``A v = T v + T^T v - diag(T) v``)."""
    keep = tensor.indices[:, 1] == tensor.indices[:, 2]
    return type(tensor)(tensor.features[keep], tensor.indices[keep], tensor.spatial_shape, tensor.batch_size)


def system_matvec(systems_tril, vector_batch: torch.Tensor) -> torch.Tensor:
    """``A v`` per batch element from the stored lower triangle ``T`` (``systems += tril(systems, -1)^T``, metrics.py:47,70)."""
    return (sparse_matvec_mul(systems_tril, vector_batch, False) + sparse_matvec_mul(systems_tril, vector_batch, True)
            - sparse_matvec_mul(_diagonal_part(systems_tril), vector_batch, False))


def hutchinson_trace(systems_tril, preconditioners_tril, vector: torch.Tensor | None = None) -> torch.Tensor:
    """``mean_b || L (L^T v) - A v ||_2`` (``metrics.py:58-77``); ``vector`` ``[B, N]`` fixes the probes (default: randn)."""
    batch, dof = systems_tril.batch_size, systems_tril.spatial_shape[0]
    if vector is None:
        vector = torch.randn(batch, dof, device=systems_tril.features.device)
    interim = sparse_matvec_mul(preconditioners_tril, sparse_matvec_mul(preconditioners_tril, vector, True), False)
    interim = interim - system_matvec(systems_tril, vector)
    return torch.linalg.vector_norm(interim, ord=2, dim=1).mean()


def inverse_loss(systems_tril, preconditioners_tril, probes: int | None = None, generator=None) -> torch.Tensor:
    """``mean_b || L L^T A - I ||_F`` (``metrics.py:34-55``) without dense matrices: ``sum_k || (L L^T A - I) v_k ||^2`` over
    unit vectors (``probes=None``: exact) or ``probes`` Rademacher vectors (scaled: unbiased for the squared norm)."""
    batch, dof = systems_tril.batch_size, systems_tril.spatial_shape[0]
    device = systems_tril.features.device if systems_tril.features.is_cuda else torch.device("cuda")
    total = torch.zeros(batch, device=device)

    def residual(v):  # (L L^T A - I) v
        w = system_matvec(systems_tril, v)
        w = sparse_matvec_mul(preconditioners_tril, sparse_matvec_mul(preconditioners_tril, w, True), False)
        return w - v

    if probes is None:
        for k in range(dof):
            v = torch.zeros(batch, dof, device=device)
            v[:, k] = 1.0
            total = total + residual(v).square().sum(dim=1)
    else:
        for _ in range(int(probes)):
            v = torch.randint(0, 2, (batch, dof), device=device, generator=generator).float() * 2 - 1
            total = total + residual(v).square().sum(dim=1) / probes
    return total.sqrt().mean()
