"""On-disk pressure-correction systems -> the hot path's input tuple (first half).

The reference stores one folder per linear system (``uibk/deep_preconditioning/generate_data.py:109-111``)::

    sludge_patterns/case_0000/matrix.npz            scipy.sparse.save_npz of the COO matrix (keys row, col, format, shape, data)
                              right_hand_side.csv   numpy.savetxt, one value per line
                              solution.csv          numpy.savetxt

and ``SludgePatternDataSet.__getitem__`` (``data_set.py:73-130``) turns a batch of folders into
``(lower-triangular SparseConvTensor, solutions [B, N_max], right_hand_sides [B, N_max], original_sizes)``: the lower
triangle of every matrix (``row >= col``), padded with trivial equations (unit diagonal, solution and right-hand side
1) up to the largest system of the whole data set, features fp32, indices int32 ``(batch, row, col)``.

:class:`SludgePatternDataSet` here reads exactly that layout and returns exactly that tuple (with the spconv-free
:class:`~deeppreconditioning_b200.model.SparseConvTensor`), so ``BenchmarkSuite(SludgePatternDataSet("test", 1, False,
root), model)`` works on a dataset produced by the reference's generator; :func:`write_case` writes the layout, which
is how the synthetic systems of :mod:`.synthetic` can be dumped for the reference's own pipeline. The dataset itself
(OpenFOAM runs) is not available offline; the format is pinned by the round-trip test in ``tests/test_host.py``.
"""

from __future__ import annotations

import random
from pathlib import Path

import numpy as np
import scipy.sparse as sp
import torch

from .model import SparseConvTensor

ROOT: Path = Path("./assets/data/raw/")  # data_set.py:20


def write_case(directory: Path, matrix, right_hand_side, solution) -> None:
    """One system in the reference's on-disk layout (generate_data.py:109-111)."""
    directory = Path(directory)
    directory.mkdir(parents=True, exist_ok=True)
    sp.save_npz(directory / "matrix.npz", sp.coo_matrix(matrix), compressed=False)
    np.savetxt(directory / "right_hand_side.csv", np.asarray(right_hand_side))
    np.savetxt(directory / "solution.csv", np.asarray(solution))


class SludgePatternDataSet:
    """``data_set.py:23-130`` without spconv: same constructor arguments, split, padding and return tuple."""

    def __init__(self, stage: str, batch_size: int, shuffle: bool = True, root: Path = ROOT, device=None) -> None:
        self._folders = sorted((Path(root) / "sludge_patterns").glob("case_*"))  # data_set.py:38
        split = len(self._folders) * 80 // 100  # data_set.py:42-44: 80/20
        if stage == "train":
            self.folders = self._folders[:split]
        elif stage == "test":
            self.folders = self._folders[split:]
        else:
            raise AssertionError(f"Invalid stage {stage}")
        if shuffle:
            random.shuffle(self.folders)  # data_set.py:48
        self.batch_size = batch_size
        self.dof_max = self._compute_max_dof()
        self.device = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")

    def _compute_max_dof(self) -> int:
        """data_set.py:56-68: the largest system of the WHOLE data set (both stages)."""
        max_dof = 0
        for folder in self._folders:
            max_dof = max(max_dof, int(np.load(folder / "matrix.npz")["shape"].max()))
        assert max_dof > 0, "Maximum degrees of freedom is zero"
        return max_dof

    def __len__(self) -> int:
        return len(self.folders) // self.batch_size

    def __getitem__(self, index: int):
        if index >= len(self):
            raise IndexError(index)
        features, indices, solutions, right_hand_sides, original_sizes = [], [], [], [], ()
        for batch_index in range(self.batch_size):
            folder = self.folders[index * self.batch_size + batch_index]
            with np.load(folder / "matrix.npz") as npz:
                rows, columns, values, shape = npz["row"], npz["col"], npz["data"], npz["shape"]
            n = int(shape[0])
            original_sizes += (n,)
            difference = self.dof_max - n
            keep = rows >= columns  # data_set.py:91-95: lower triangle, the matrix is symmetric
            rows, columns, values = rows[keep], columns[keep], values[keep]
            pad = np.arange(n, self.dof_max)  # data_set.py:96-99: trivial equations up to dof_max
            rows, columns = np.append(rows, pad), np.append(columns, pad)
            values = np.append(values, np.ones(difference))
            solution = np.loadtxt(folder / "solution.csv", ndmin=1)
            rhs = np.loadtxt(folder / "right_hand_side.csv", ndmin=1)
            features.append(values[:, None])
            indices.append(np.column_stack((np.full(len(values), batch_index), rows, columns)))
            solutions.append(np.pad(solution, (0, difference), constant_values=1)[None, :])
            right_hand_sides.append(np.pad(rhs, (0, difference), constant_values=1)[None, :])
        to = lambda a, dtype: torch.from_numpy(np.vstack(a)).to(dtype).to(self.device)
        tril = SparseConvTensor(to(features, torch.float32), to(indices, torch.int32), [self.dof_max, self.dof_max],
                                self.batch_size)
        return tril, to(solutions, torch.float32), to(right_hand_sides, torch.float32), original_sizes
