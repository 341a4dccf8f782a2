"""Drop-in for the hot-path part of ``uibk/deep_preconditioning/test.py``: ``BenchmarkSuite``.

Same class name, constructor fields, private helper names and ``run()`` / ``dump_csv()`` flow as the reference
(``test.py:31-198``); what changes is what the helpers return and where the work runs:

* ``_reconstruct_system``  (test.py:61-68)   -> device :class:`CsrMatrix` via the K1 ``symmetrise`` kernel, not a dense N x N tensor;
* ``_construct_learned``   (test.py:100-105) -> :class:`FactoredMultiply` on the CSR of ``L`` and ``L^T`` (no dense fp32 ``L @ L.T``);
* ``_construct_incomplete_cholesky`` (test.py:81-88) -> IC(0) applied by triangular solves (the reference
  multiplies by ``L L^T``, which approximates ``A`` instead of ``A^-1``, and tags the technique "unstable");
* ``preconditioned_conjugate_gradient`` (test.py:138) -> the fused B200 solve.

* ``_compute_kappa`` (test.py:111-113, a dense ``cond(M @ A)``) -> the Lanczos estimate ``lambda_max / lambda_min`` of the
  preconditioned operator from the solve's own CG coefficients (:mod:`.spectrum`).

Out of scope: dense singular values (``_compute_eigenvalues``, O(n^3)), histogram plots, the ``main()``
pipeline glue (dvc, checkpoint loading).
"""

from __future__ import annotations

import time
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np
import torch

from .cg import PcgBatch
from .precond import FactoredMultiply, FactoredSolve, Identity, Jacobi, incomplete_cholesky0, level_ordering
from .sparse import CsrMatrix

RESULTS_DIRECTORY: Path = Path("./assets/results/")


@dataclass
class BenchmarkSuite:
    """Preconditioner benchmark suite (test.py:31-41).

    Args:
        data_set: The test data set to benchmark on; items are ``(systems_tril, solutions, rhs, original_sizes)``.
        model: The fully convolutional model that maps ``tril(A)`` to ``L``.
    """

    data_set: object
    model: torch.nn.Module
    techniques: tuple[str, ...] = ("vanilla", "jacobi", "incomplete_cholesky", "learned")
    rtol: float = 1e-8       # cg.py:51
    max_iter: int = 1024     # cg.py:51
    level_order_ic: bool = False  # solve the IC(0) comparator on the system renumbered by the factor's level sets
    kappas: dict = field(default_factory=dict)
    densities: dict = field(default_factory=dict)
    iterations: dict = field(default_factory=dict)
    setups: dict = field(default_factory=dict)
    durations: dict = field(default_factory=dict)
    totals: dict = field(default_factory=dict)
    successes: dict = field(default_factory=dict)
    residuals: dict = field(default_factory=dict)

    def __post_init__(self) -> None:
        for store in (self.kappas, self.densities, self.iterations, self.setups, self.durations, self.totals,
                      self.successes, self.residuals):
            for name in self.techniques:
                store.setdefault(name, [])

    # ---- operand construction ----------------------------------------------------------------------------------
    def _reconstruct_system(self, system_tril, original_size: int) -> CsrMatrix:
        """Reconstruct the linear system from the sparse lower-triangular tensor (test.py:61-68)."""
        assert system_tril.batch_size == 1, "Set batch size to one for testing"
        return CsrMatrix.from_spconv(system_tril, original_size, mode="symmetrise")

    def _construct_vanilla(self, matrix: CsrMatrix):
        """The baseline which is no preconditioner (test.py:70-72)."""
        return Identity()

    def _construct_jacobi(self, matrix: CsrMatrix):
        """The Jacobi preconditioner (test.py:74-79)."""
        return Jacobi(matrix)

    def _construct_incomplete_cholesky(self, matrix: CsrMatrix, fill_in: int = 0, threshold: float = 0.0):
        """The incomplete Cholesky preconditioner (test.py:81-88), IC(0) only (``ilupp.ichol0`` branch)."""
        if fill_in != 0 or threshold != 0.0:
            raise NotImplementedError("only IC(0) (fill_in=0, threshold=0.0) is implemented on the GPU")
        return FactoredSolve(incomplete_cholesky0(self._tril_of(matrix)))

    def _construct_learned(self, system_tril, original_size: int):
        """Our preconditioner (test.py:100-105): model forward, then CSR of ``L`` and ``L^T`` on the device."""
        with torch.no_grad():
            preconditioners_tril = self.model(system_tril)
        lower = CsrMatrix.from_spconv(preconditioners_tril, original_size, mode="tril")
        lower_t = CsrMatrix.from_spconv(preconditioners_tril, original_size, mode="tril_t")
        return FactoredMultiply(lower, lower_t)

    def _tril_of(self, matrix: CsrMatrix) -> CsrMatrix:
        return self._current_tril if getattr(self, "_current_tril", None) is not None else matrix

    def _compute_sparsity(self, preconditioner) -> float:
        """Density in percent of the explicit ``M`` the reference would store (test.py:107-109)."""
        if isinstance(preconditioner, Identity):
            n = self._current_n
            return 100 * n / (n * n)
        if isinstance(preconditioner, Jacobi):
            n = preconditioner.dinv.shape[0]
            return 100 * n / (n * n)
        lower = preconditioner.L  # symbolic nnz(L L^T) on the host: a diagnostic, not on the timed path
        import scipy.sparse as sp

        rowptr, col, _ = lower.to_host()
        pattern = sp.csr_matrix((np.ones(len(col), np.int8), col, rowptr), shape=lower.shape)
        return 100 * (pattern @ pattern.T).nnz / (lower.n * lower.n)

    # ---- the benchmark loop ----------------------------------------------------------------------------------------
    def run(self) -> None:
        """Run the whole benchmark suite (test.py:119-149)."""
        for index in range(len(self.data_set)):
            system_tril, _, right_hand_side, original_size = self.data_set[index]
            n = int(original_size[0])
            device = torch.device("cuda", torch.cuda.current_device())
            if not system_tril.indices.is_cuda:
                system_tril = type(system_tril)(system_tril.features.to(device), system_tril.indices.to(device),
                                                system_tril.spatial_shape, system_tril.batch_size)
            matrix = self._reconstruct_system(system_tril, n)
            self._current_n = n
            self._current_tril = CsrMatrix.from_spconv(system_tril, n, mode="tril")
            rhs = right_hand_side[0, :n].squeeze().to(device=device, dtype=torch.float64)  # test.py:124

            for name in self.techniques:
                torch.cuda.synchronize()
                start_time = time.perf_counter()
                if name == "learned":
                    preconditioner = self._construct_learned(system_tril, n)
                else:
                    preconditioner = getattr(self, f"_construct_{name}")(matrix)
                torch.cuda.synchronize()
                setup = time.perf_counter() - start_time if name != "vanilla" else 0.0  # test.py:135

                system, b_sys = matrix, rhs
                if name == "incomplete_cholesky" and self.level_order_ic:
                    # same solve, rows renumbered so that a level of the factor is a contiguous run (precond.LevelOrdering):
                    # the triangular solves read the factor and the vectors coalesced. Counted as set-up time.
                    torch.cuda.synchronize()
                    start_time = time.perf_counter()
                    order = level_ordering(self._current_tril)
                    renumbered = order.renumber(system_tril)
                    system = CsrMatrix.from_spconv(renumbered, n, mode="symmetrise")
                    preconditioner = FactoredSolve(incomplete_cholesky0(CsrMatrix.from_spconv(renumbered, n, mode="tril")))
                    b_sys = order.to_level(rhs)
                    torch.cuda.synchronize()
                    setup += time.perf_counter() - start_time

                density = self._compute_sparsity(preconditioner)
                batch = PcgBatch([(system, b_sys, preconditioner)], self.rtol, self.max_iter, history=True)
                torch.cuda.synchronize()
                start_time = time.perf_counter()
                batch.solve()
                torch.cuda.synchronize()
                duration = time.perf_counter() - start_time
                result = batch.results(duration)[0]

                # test.py:111-113 takes a dense cond(M @ A); here the solve's own CG coefficients give the Lanczos
                # estimate lambda_max / lambda_min of the preconditioned operator (spectrum.py)
                self.kappas[name].append(result.kappa)
                self.densities[name].append(density)
                self.iterations[name].append(result.iterations)
                self.setups[name].append(setup)
                self.durations[name].append(duration)
                self.totals[name].append(setup + duration)
                self.successes[name].append(100 * (1 - result.info))
                self.residuals[name].append(result.res)

    def dump_csv(self, directory: Path | None = None) -> None:
        """Dump the durations and iterations to CSV files in the reference's layout (test.py:175-198)."""
        directory = Path(directory) if directory is not None else RESULTS_DIRECTORY
        directory.mkdir(parents=True, exist_ok=True)
        parameters = ["kappas", "densities", "iterations", "setups", "durations", "totals", "successes"]
        with (directory / "table.csv").open(mode="w") as file_io:
            file_io.write("technique," + ",".join(parameters) + "\n")
            for technique in self.techniques:
                line = technique
                for parameter in parameters:
                    line += "," + str(np.mean(getattr(self, parameter)[technique], dtype=float))
                file_io.write(line + "\n")
        with (directory / "totals.csv").open(mode="w") as file_io:
            file_io.write(",".join(self.techniques) + "\n")
            for index in range(len(self.totals[self.techniques[0]])):
                file_io.write(",".join(str(self.totals[t][index]) for t in self.techniques) + "\n")
