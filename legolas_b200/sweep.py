"""Sharding of independent eigenproblems over the GPUs of one node.

One eigenproblem does not shard (the factorisation is a 1-D recurrence, Arnoldi needs global
reductions every step); what shards naturally is a list of independent units — shifts sigma
for one equilibrium (multi-shift spectrum scan) or (k2, k3) wavenumber points (parameter
sweep).  The reference runs such sweeps as independent OS processes, one per parfile
(post_processing/pylbo/automation/runner.py:202-211); here it is one process per GPU
(torch.distributed), a static round-robin partition of the unit list, no data-path
collective, and one gather of the eigenvalues at the end.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np


def partition(n_units: int, rank: int, world_size: int) -> List[int]:
    """Round-robin unit indices of ``rank`` (unit i goes to rank i % world_size)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    return list(range(rank, n_units, world_size))


def gather_eigenvalues(local: np.ndarray, local_ids: Sequence[int], n_units: int, nev: int,
                       group=None) -> np.ndarray:
    """All ranks contribute ``local`` (len(local_ids) x nev complex, NaN-padded) and every rank
    gets the full (n_units x nev) table.  Works with the nccl backend (device tensors) and with
    gloo (host tensors); without torch.distributed initialised it is a local scatter."""
    import torch
    import torch.distributed as dist

    table = np.full((n_units, nev), np.nan + 1j * np.nan, dtype=np.complex128)
    local = np.asarray(local, dtype=np.complex128).reshape(len(local_ids), nev)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        for row, uid in zip(local, local_ids):
            table[uid] = row
        return table
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per_rank = (n_units + world - 1) // world
    use_cuda = dist.get_backend(group) == "nccl"
    device = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    # fixed-size payload per rank: [unit id, re/im of nev eigenvalues] per slot
    payload = torch.full((per_rank, 1 + 2 * nev), float("nan"), dtype=torch.float64)
    for slot, (row, uid) in enumerate(zip(local, local_ids)):
        payload[slot, 0] = float(uid)
        payload[slot, 1::2] = torch.from_numpy(row.real.copy())
        payload[slot, 2::2] = torch.from_numpy(row.imag.copy())
    payload = payload.to(device)
    out = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(out, payload, group=group)
    for chunk in out:
        chunk = chunk.cpu().numpy()
        for slot in range(per_rank):
            if np.isnan(chunk[slot, 0]):
                continue
            table[int(chunk[slot, 0])] = chunk[slot, 1::2] + 1j * chunk[slot, 2::2]
    del rank
    return table


def run_sweep(units: Sequence, solve_unit: Callable, nev: int, rank: int = 0, world_size: int = 1,
              group=None) -> np.ndarray:
    """Solve ``units[i]`` on rank ``i % world_size`` with ``solve_unit(unit) -> omega (nev)`` and
    gather the eigenvalue table on every rank."""
    ids = partition(len(units), rank, world_size)
    local = np.full((len(ids), nev), np.nan + 1j * np.nan, dtype=np.complex128)
    for slot, uid in enumerate(ids):
        omega = np.asarray(solve_unit(units[uid]), dtype=np.complex128)
        local[slot, :len(omega)] = omega[:nev]
    return gather_eigenvalues(local, ids, len(units), nev, group=group)
