"""Synthetic pressure-correction systems in the reference's input layout.

The OpenFOAM dataset of the reference is not available offline, so inputs are synthetic SPD
variable-coefficient finite-volume Laplacians that mimic ``fvm::laplacian(rAUf, p_rgh)``
(reference ``foam/newInterFoam/pEqn.H:45``) after the sign flip of
``uibk/deep_preconditioning/generate_data.py:71``:

* face coefficients ``k_f = exp(sigma * N(0, 1))`` on every face, boundary faces included
  (Dirichlet folded into the diagonal, the analogue of ``setReference``, ``pEqn.H:48``);
* off-diagonals ``-k_f``, diagonal ``sum`` of the adjacent ``k_f``; natural (lexicographic) ordering;
* generated in fp64 and **rounded to fp32**, because the reference stores matrix values as fp32
  features (``data_set.py:121``) and widens them back to fp64 for the solve (``test.py:68``);
* right-hand side ``rng.uniform(-1, 1, N)`` (``generate_data.py:106``), fp32-rounded (``data_set.py:128``);
* ``numpy.random.default_rng(69420 + system_index)`` (``generate_data.py:15``).

Everything is returned the way ``SludgePatternDataSet.__getitem__`` (``data_set.py:73-130``) returns
it: a lower-triangular ``SparseConvTensor`` (features ``[nnz,1]`` fp32, indices ``[nnz,3]`` int32
``(batch,row,col)``), solutions, right-hand sides ``[B,N]`` fp32 and the tuple of original sizes.
"""

from __future__ import annotations

import numpy as np
import torch

from .model import SparseConvTensor

SEED_BASE = 69420  # generate_data.py:15


def _rng(system_index: int) -> np.random.Generator:
    return np.random.default_rng(SEED_BASE + int(system_index))


def poisson2d_tril(n: int, sigma: float = 0.5, system_index: int = 0):
    """Lower triangle (diagonal included) of an ``n x n``-cell 5-point FV Laplacian.

    Returns ``(rows int32, cols int32, vals float32, rhs float32)`` with ``N = n*n`` unknowns,
    cell ``(j, i)`` -> unknown ``i + n*j``. Entries come diagonal first, then the x-neighbours, then
    the y-neighbours: deliberately *not* sorted by (row, col), like the reference's COO files.
    """
    rng = _rng(system_index)
    kx = np.exp(sigma * rng.standard_normal((n, n + 1)))  # x-faces of row j: i = 0..n
    ky = np.exp(sigma * rng.standard_normal((n + 1, n)))  # y-faces of column i: j = 0..n
    diag = kx[:, :-1] + kx[:, 1:] + ky[:-1, :] + ky[1:, :]
    idx = (np.arange(n)[None, :] + n * np.arange(n)[:, None]).astype(np.int64)

    rows = [idx.ravel(), idx[:, 1:].ravel(), idx[1:, :].ravel()]
    cols = [idx.ravel(), idx[:, :-1].ravel(), idx[:-1, :].ravel()]
    vals = [diag.ravel(), -kx[:, 1:-1].ravel(), -ky[1:-1, :].ravel()]
    rhs = rng.uniform(-1.0, 1.0, n * n)
    return (
        np.concatenate(rows).astype(np.int32),
        np.concatenate(cols).astype(np.int32),
        np.concatenate(vals).astype(np.float32),
        rhs.astype(np.float32),
    )


def poisson3d_tril(n: int, sigma: float = 0.5, system_index: int = 0):
    """Lower triangle of an ``n^3``-cell 7-point FV Laplacian, unknown ``i + n*j + n*n*k``."""
    rng = _rng(system_index)
    kx = np.exp(sigma * rng.standard_normal((n, n, n + 1)))
    ky = np.exp(sigma * rng.standard_normal((n, n + 1, n)))
    kz = np.exp(sigma * rng.standard_normal((n + 1, n, n)))
    diag = kx[:, :, :-1] + kx[:, :, 1:] + ky[:, :-1, :] + ky[:, 1:, :] + kz[:-1] + kz[1:]
    i = np.arange(n, dtype=np.int64)
    idx = i[None, None, :] + n * i[None, :, None] + n * n * i[:, None, None]

    rows = [idx.ravel(), idx[:, :, 1:].ravel(), idx[:, 1:, :].ravel(), idx[1:].ravel()]
    cols = [idx.ravel(), idx[:, :, :-1].ravel(), idx[:, :-1, :].ravel(), idx[:-1].ravel()]
    vals = [diag.ravel(), -kx[:, :, 1:-1].ravel(), -ky[:, 1:-1, :].ravel(), -kz[1:-1].ravel()]
    rhs = rng.uniform(-1.0, 1.0, n**3)
    return (
        np.concatenate(rows).astype(np.int32),
        np.concatenate(cols).astype(np.int32),
        np.concatenate(vals).astype(np.float32),
        rhs.astype(np.float32),
    )


def make_batch(kind: str, n: int, system_indices, sigma: float = 0.5, device="cpu", pad_to: int | None = None):
    """Build one reference-style batch ``(systems_tril, solutions, right_hand_sides, original_sizes)``.

    Mirrors ``SludgePatternDataSet.__getitem__`` (``data_set.py:73-130``): every system is padded with
    trivial equations (unit diagonal, rhs 1) up to ``pad_to`` unknowns when given (``data_set.py:95-118``).
    ``solutions`` is a zero placeholder (the reference stores scipy-CG solutions there; the PCG path
    never reads it, ``test.py:122``).
    """
    gen = {"poisson2d": poisson2d_tril, "poisson3d": poisson3d_tril}[kind]
    feats, inds, rhss, sizes = [], [], [], []
    for b, s in enumerate(system_indices):
        r, c, v, rhs = gen(n, sigma, s)
        size = rhs.shape[0]
        dof = size if pad_to is None else int(pad_to)
        assert dof >= size
        if dof > size:
            extra = np.arange(size, dof, dtype=np.int32)
            r, c = np.concatenate([r, extra]), np.concatenate([c, extra])
            v = np.concatenate([v, np.ones(dof - size, np.float32)])
            rhs = np.concatenate([rhs, np.ones(dof - size, np.float32)])
        feats.append(v[:, None])
        inds.append(np.column_stack([np.full(len(v), b, np.int32), r, c]))
        rhss.append(rhs[None, :])
        sizes.append(size)
    dof = rhss[0].shape[1]
    assert all(x.shape[1] == dof for x in rhss), "pad_to is required for ragged batches"
    features = torch.from_numpy(np.vstack(feats)).float().to(device)
    indices = torch.from_numpy(np.vstack(inds)).int().to(device)
    systems_tril = SparseConvTensor(features, indices, [dof, dof], len(sizes))
    right_hand_sides = torch.from_numpy(np.vstack(rhss)).float().to(device)
    solutions = torch.zeros_like(right_hand_sides)
    return systems_tril, solutions, right_hand_sides, tuple(sizes)


class SyntheticPressureDataSet:
    """Drop-in for ``SludgePatternDataSet(stage, batch_size, shuffle)`` (``data_set.py:23-130``) on synthetic systems."""

    def __init__(self, kind: str = "poisson2d", n: int = 64, number_samples: int = 4, batch_size: int = 1,
                 sigma: float = 0.5, device="cpu", first_index: int = 0) -> None:
        self.kind, self.n, self.sigma, self.device = kind, n, sigma, device
        self.batch_size = batch_size
        self.indices = list(range(first_index, first_index + number_samples))

    def __len__(self) -> int:
        return len(self.indices) // self.batch_size

    def __getitem__(self, index: int):
        if index >= len(self):
            raise IndexError(index)
        chunk = self.indices[index * self.batch_size:(index + 1) * self.batch_size]
        return make_batch(self.kind, self.n, chunk, self.sigma, self.device)
