"""B200-native PCG hot path of jsappl/DeepPreconditioning.

Python/PyTorch host code with the reference's call signatures (``cg.py``, ``test.py``, ``utils.py``) over a C-ABI
library of hand-written sm_100a CUDA kernels (``include/dpcg.h``, ``deeppreconditioning_b200/csrc``). Importing the
package does not load the library; the first call that needs the GPU does, and raises if it is missing.
"""

from .cg import (PcgBatch, PcgResult, conjugate_gradient, pcg_solve, pcg_solve_batch,  # noqa: F401
                 preconditioned_conjugate_gradient, stopping_criterion)
from .precond import (CsrOperator, FactoredMultiply, FactoredSolve, Identity, Jacobi, analyse,  # noqa: F401
                      incomplete_cholesky0, triangular_solve)
from .sparse import CsrMatrix, as_csr  # noqa: F401

__version__ = "0.1.0"
