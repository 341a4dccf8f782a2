"""Sharding of the independent test-set systems over the GPUs of one box.

The reference evaluates its test set in a serial Python loop (``uibk/deep_preconditioning/test.py:121``); systems share
no state except appended result lists (``test.py:143-149``). Rank ``r`` of ``W`` takes systems ``r, r+W, ...``
(interleaved, so slow and fast systems mix), solves them with the single-GPU batched path, and ONE collective at the
end gathers a fixed-size record per system. No collective runs inside a solve.
"""

from __future__ import annotations

import torch
import torch.distributed as dist

RECORD_WIDTH = 4  # (system index, iterations, criterion value, milliseconds)


def shard_indices(n_systems: int, rank: int, world_size: int) -> list[int]:
    """Interleaved partition: counts differ by at most one across ranks."""
    return list(range(rank, n_systems, world_size))


def make_records(indices, results, milliseconds) -> torch.Tensor:
    """``[n_local, 4]`` fp64 records from :class:`~deeppreconditioning_b200.cg.PcgResult` objects."""
    rows = [[float(i), float(r.iterations), float(r.res), float(ms)] for i, r, ms in zip(indices, results, milliseconds)]
    return torch.tensor(rows, dtype=torch.float64).reshape(-1, RECORD_WIDTH)


def gather_records(local: torch.Tensor, n_systems: int, group=None) -> torch.Tensor:
    """All-gather the per-system records and return them ordered by system index: ``[n_systems, width]`` on every rank
    (column 0 is the system index; ``bench.py`` uses the 4 columns of :func:`make_records`, ``BenchmarkSuite.run`` 8).

    Uses the process group's backend (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests); the payload is
    ``n_systems * 32`` bytes — latency-only. Works without an initialised process group (single process).
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        gathered = local
    else:
        world = dist.get_world_size(group)
        per_rank = (n_systems + world - 1) // world
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        width = local.shape[1]
        padded = torch.full((per_rank, width), -1.0, dtype=torch.float64, device=device)
        padded[: local.shape[0]] = local.to(device)
        out = torch.empty((world * per_rank, width), dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(out, padded, group=group)
        gathered = out[out[:, 0] >= 0].cpu()
    order = torch.argsort(gathered[:, 0])
    gathered = gathered[order]
    assert gathered.shape[0] == n_systems, f"gathered {gathered.shape[0]} records for {n_systems} systems"
    return gathered
