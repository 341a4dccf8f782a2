"""Drop-in for ``uibk/deep_preconditioning/cg.py`` backed by the B200 kernels of ``libdpcg``.

Same names, argument meaning and return conventions as the reference:

* ``stopping_criterion(_, rk, b)``                                   cg.py:15-17
* ``conjugate_gradient(A, b, x0, x_true, rtol, max_iter)``           cg.py:20-47  -> ``(errors, x_hat)``
* ``preconditioned_conjugate_gradient(A, b, M, x0, x_true, rtol, max_iter)``  cg.py:50-90
  -> ``(seconds, iterations, info)`` with ``info`` always 0 and no exception on non-convergence
  (``iterations == max_iter``), inputs never mutated.

``A`` may be a :class:`~deeppreconditioning_b200.sparse.CsrMatrix`, a dense/sparse torch tensor on any device (what
``test.py:138`` and ``train.py:102`` pass) or a scipy matrix; ``M`` one of the operators in
:mod:`~deeppreconditioning_b200.precond`, ``None``, or any explicit matrix. Host operands are copied to the GPU — the
arithmetic never runs on the CPU. :func:`pcg_solve` / :func:`pcg_solve_batch` additionally return what the reference
computes but drops (``x_hat``, last criterion value, history).
"""

from __future__ import annotations

import ctypes
import time
from dataclasses import dataclass, field

import torch

from . import _lib
from .precond import Identity, _Operator, as_operator, fill_packed
from .sparse import CsrMatrix, _workspace, as_csr, pack_many

_ENGINES = {"fused": _lib.ENGINE_FUSED, "stepped": _lib.ENGINE_STEPPED}


@dataclass
class PcgResult:
    seconds: float
    iterations: int
    info: int
    x_hat: torch.Tensor
    res: float
    history: list = field(default_factory=list)
    alphas: list = field(default_factory=list)  # a of every body (cg.py:78), with history=True
    betas: list = field(default_factory=list)   # beta behind every body's p (cg.py:82; betas[0] = 0)

    @property
    def kappa(self) -> float:
        """Condition-number estimate of ``M A`` from the CG coefficients (see :mod:`.spectrum`); needs history=True."""
        from .spectrum import kappa_estimate

        return kappa_estimate(self.alphas, self.betas) if self.alphas else float("nan")


def stopping_criterion(_, rk, b):
    """Squared relative residual (cg.py:15-17)."""
    return torch.inner(rk, rk) / torch.inner(b, b)


class PcgBatch:
    """A prepared batch of independent systems: operands resident on one GPU, scratch allocated, ready to launch.

    ``solve()`` enqueues ONE call of ``dp_pcg_solve_f64`` for the whole batch; ``results()`` reads iterations,
    criterion and solutions back. Preparation (``__init__``) is the analogue of the reference's ``setup`` time
    (``test.py:130-135``), ``solve()`` of its ``duration`` (``cg.py:69-88``).
    """

    def __init__(self, systems, rtol: float = 1e-8, max_iter: int = 1024, engine: str = "fused",
                 check_every: int = 32, history: bool = False, device=None, pack: bool = True) -> None:
        """``pack``: give every matrix the batch streams its lossless packed copy (``dp_csr_pack``; 6 instead of 12 bytes
        per entry, bit-identical results) and let the fused engine stream those when all of them exist."""
        if not systems:
            raise ValueError("empty batch")
        self.device = torch.device(device) if device is not None else None
        self.rtol, self.max_iter = float(rtol), int(max_iter)
        self.params = _lib.PcgParams(self.rtol, self.max_iter, _ENGINES[engine], int(check_every), 0)
        self.entries = []
        lib = _lib.lib()
        for item in systems:
            A, b, M = item[0], item[1], item[2]
            x0 = item[3] if len(item) > 3 else None
            A = as_csr(A, self.device)
            if self.device is None:
                self.device = A.device
            if A.device != self.device:
                raise _lib.DpcgError("all systems of a batch must live on one device")
            M = as_operator(M, self.device)
            n = A.n
            b_dev = b.detach().to(device=self.device, dtype=torch.float64).contiguous()
            if b_dev.shape != (n,):
                raise ValueError(f"right-hand side has shape {tuple(b_dev.shape)}, expected ({n},)")
            f64 = dict(dtype=torch.float64, device=self.device)
            x = (x0.detach().to(**f64).clone() if x0 is not None else torch.zeros(n, **f64)).contiguous()
            work = torch.empty(lib.dp_pcg_work_doubles(n), **f64)
            hist = torch.full((self.max_iter + 1,), float("nan"), **f64) if history else None
            coef = torch.full((2 * (self.max_iter + 1),), float("nan"), **f64) if history else None
            self.entries.append(dict(A=A, M=M, b=b_dev, x=x, work=work, hist=hist, coef=coef, out_device=b.device))
        nsys = len(self.entries)
        if any(getattr(e["M"], "tile_stream", False) for e in self.entries):
            # tile-stream solves are separate launches between the phases: a feature of the stepped engine. Converged
            # systems drop out of every launch on the device (state flags), so an iteration enqueued past the batch's
            # convergence is a handful of empty launches and the host polls the done counter only every 8 iterations.
            self.params = _lib.PcgParams(self.rtol, self.max_iter, _ENGINES["stepped"], min(int(check_every), 8), 0)
        solve_mode = any(e["M"].precond == _lib.PRECOND_SOLVE for e in self.entries)
        if pack and engine == "fused" and not solve_mode:  # one pass over each matrix, one synchronisation for the batch
            pack_many([m for e in self.entries for m in [e["A"], *e["M"].stream_matrices()]])
        self.iters = torch.full((nsys,), -1, dtype=torch.int32, device=self.device)
        self.res = torch.full((nsys,), float("nan"), dtype=torch.float64, device=self.device)
        self.flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.ws = _workspace(lib.dp_pcg_workspace_bytes(nsys), self.device)
        self.descs = (_lib.PcgSystem * nsys)()
        for i, e in enumerate(self.entries):
            d, A = self.descs[i], e["A"]
            d.n, d.a_nnz = A.n, A.nnz
            d.a_rowptr, d.a_col, d.a_val = _lib.ptr(A.rowptr), _lib.ptr(A.col), _lib.ptr(A.val)
            fill_packed(d, "a", A)
            e["M"].fill(d)
            if not pack:
                for tag in ("a", "m", "mt"):
                    for name in ("col16", "val32", "tile_base"):
                        setattr(d, f"{tag}_{name}", None)
            d.b, d.x, d.work = _lib.ptr(e["b"]), _lib.ptr(e["x"]), _lib.ptr(e["work"])
            d.iters_out = self.iters.data_ptr() + 4 * i
            d.res_out = self.res.data_ptr() + 8 * i
            d.history = _lib.ptr(e["hist"])
            d.coef = _lib.ptr(e["coef"])

    def __len__(self) -> int:
        return len(self.entries)

    def reset(self, x0s=None) -> None:
        """Restore the initial guess (zeros unless given) so the same prepared batch can be solved again."""
        for i, e in enumerate(self.entries):
            if x0s is not None and x0s[i] is not None:
                e["x"].copy_(x0s[i])
            else:
                e["x"].zero_()
        self.iters.fill_(-1)
        self.res.fill_(float("nan"))

    def solve(self) -> None:
        """Enqueue the solve on the current stream (asynchronous for the fused engine)."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dp_pcg_solve_f64(self.descs, len(self.entries), ctypes.byref(self.params),
                                                   _lib.ptr(self.flag), _lib.ptr(self.ws), self.ws.numel(),
                                                   _lib.stream_ptr(self.device)), "dp_pcg_solve_f64")

    def results(self, seconds: float = 0.0) -> list[PcgResult]:
        """Synchronise and collect ``PcgResult`` per system (``x_hat`` on the device the right-hand side came from)."""
        _lib.raise_on_flag(self.flag, "dp_pcg_solve_f64")
        iters, res = self.iters.cpu().tolist(), self.res.cpu().tolist()
        # solutions that go back to the host leave through pinned buffers, all copies in flight before one wait
        xs = []
        for e in self.entries:
            if e["out_device"].type == "cpu":
                host = torch.empty(e["x"].shape, dtype=torch.float64, pin_memory=True)
                host.copy_(e["x"], non_blocking=True)
                xs.append(host)
            else:
                xs.append(e["x"].to(e["out_device"]))
        torch.cuda.synchronize(self.device)
        out = []
        for i, e in enumerate(self.entries):
            hist = e["hist"][: iters[i] + 1].cpu().tolist() if e["hist"] is not None and iters[i] >= 0 else []
            alphas, betas = [], []
            if e["coef"] is not None and iters[i] > 0:
                c = e["coef"][: 2 * iters[i]].cpu().view(-1, 2)
                alphas, betas = c[:, 0].tolist(), c[:, 1].tolist()
                betas[0] = 0.0
            out.append(PcgResult(seconds, int(iters[i]), 0, xs[i], float(res[i]), hist, alphas, betas))
        return out


def pcg_solve_batch(systems, rtol: float = 1e-8, max_iter: int = 1024, engine: str = "fused", check_every: int = 32,
                    history: bool = False, device=None, pack: bool = True) -> list[PcgResult]:
    """Solve independent systems ``[(A, b, M[, x0]), ...]`` in one launch; ``seconds`` is the batch wall time."""
    batch = PcgBatch(systems, rtol, max_iter, engine, check_every, history, device, pack)
    torch.cuda.synchronize(batch.device)
    start = time.perf_counter()
    batch.solve()
    torch.cuda.synchronize(batch.device)
    return batch.results(time.perf_counter() - start)


def pcg_solve(A, b, M=None, x0=None, rtol: float = 1e-8, max_iter: int = 1024, engine: str = "fused",
              check_every: int = 32, history: bool = False, pack: bool = True) -> PcgResult:
    """One system; like :func:`preconditioned_conjugate_gradient` but returns the full :class:`PcgResult`."""
    return pcg_solve_batch([(A, b, M, x0)], rtol, max_iter, engine, check_every, history, pack=pack)[0]


def preconditioned_conjugate_gradient(A, b: torch.Tensor, M, x0=None, x_true=None, rtol=1e-8, max_iter=1024):
    """The preconditioned conjugate gradient method (cg.py:50-90): returns ``(seconds, iterations, 0)``.

    ``seconds`` spans the solve only, like the reference's ``perf_counter`` pair (cg.py:69,88); operand upload and
    workspace allocation are outside it. ``x_true`` is accepted for signature parity; the A-norm error it feeds in the
    reference (cg.py:65,85,87) is computed and dropped there, so it is not evaluated here.
    """
    result = pcg_solve(A, b, M, x0, rtol, max_iter)
    return result.seconds, result.iterations, 0


def conjugate_gradient(A, b, x0=None, x_true=None, rtol=1e-8, max_iter=1024):
    """Unpreconditioned CG (cg.py:20-47): returns ``(errors, x_hat)``, ``errors`` a list of ``(A-norm error, res)``.

    With ``M = I`` the preconditioned recurrence is the same arithmetic as cg.py:31-45 (``z = r``, so ``<r,z> = <r,r>``).
    The A-norm error column (``<e_k, A e_k>`` with ``e_k = x_k - x_true``, cg.py:28-30,42-44; zero without ``x_true``) is
    evaluated on the device for the final iterate (one SpMV) and carried back through the iterations by the CG identity
    ``||e_k||_A^2 - ||e_{k+1}||_A^2 = a_k <r_k, r_k>`` (Hestenes-Stiefel; numerically stable in finite precision, Strakos &
    Tichy 2002) from the step lengths and residual norms the solve records - no iterate history, no SpMV per iteration.
    """
    result = pcg_solve(A, b, Identity(), x0, rtol, max_iter, history=True)
    f64 = torch.float64
    tail = torch.zeros(len(result.history), dtype=f64)
    if x_true is not None and result.history:
        A_dev = as_csr(A)
        e = result.x_hat.to(device=A_dev.device, dtype=f64) - x_true.to(device=A_dev.device, dtype=f64)
        last = float(torch.inner(e, A_dev.matvec(e)))
        b64 = b.detach().to(f64)
        bb = float(torch.inner(b64, b64))
        drops = torch.tensor([a * r * bb for a, r in zip(result.alphas, result.history[:-1])], dtype=f64)  # a_k <r_k, r_k>
        tail[:-1] = torch.flip(torch.cumsum(torch.flip(drops, [0]), 0), [0])
        tail = tail + last
    errors = [(tail[k], torch.tensor(r, dtype=f64)) for k, r in enumerate(result.history)]
    return errors, result.x_hat
