"""Sharding of independent eigenproblems over the GPUs of one node.

One eigenproblem does not shard (the factorisation is a 1-D recurrence, Arnoldi needs global
reductions every step); what shards naturally is a list of independent units — shifts sigma
for one equilibrium (multi-shift spectrum scan) or (k2, k3) wavenumber points (parameter
sweep).  The reference runs such sweeps as independent OS processes, one per parfile, handed
out by a process pool (post_processing/pylbo/automation/runner.py:202-211); here it is one
process per GPU (torch.distributed), no data-path collective, and one gather of the
eigenvalues at the end.  Two ways to hand out the units:

* ``run_sweep``: static round-robin partition (unit i on rank i % world_size);
* ``run_queue``: a shared queue, the analogue of the reference's ``multiprocessing.Pool.imap``: the
  ranks draw the next unit index from a counter in the torch.distributed store (one atomic add
  per unit, nothing else crosses ranks), in the order given — longest first when the caller has
  a cost estimate — so that unequal units do not leave GPUs idle.  Each rank may keep several
  units in flight (``workers`` host threads, one library context and CUDA stream each): the
  kernels of a small problem leave most of a B200 idle, the units of a sweep fill it.
"""
from __future__ import annotations

import itertools
import threading
from typing import Callable, List, Optional, Sequence

import numpy as np

_QUEUE_SEQ = itertools.count()


def partition(n_units: int, rank: int, world_size: int) -> List[int]:
    """Round-robin unit indices of ``rank`` (unit i goes to rank i % world_size)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    return list(range(rank, n_units, world_size))


def _dist():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist
    return None


class WorkQueue:
    """Positions 0 .. n-1 handed out once each across all ranks (and all threads of a rank).

    With torch.distributed initialised the counter lives in the default store (TCPStore on rank 0;
    ``store.add`` is atomic); otherwise it is a local counter.  Every rank must construct its queues
    in the same order (the key is numbered per process).  A draw takes a CHUNK of consecutive positions
    per store round trip - guided self-scheduling: about 1 / (4 x drawers) of what is left, at least one -
    so that short units do not pay a round trip each while the tail of the queue is still handed out one
    by one; the positions of a chunk are served to the threads of the rank in order."""

    def __init__(self, n: int, group=None, drawers: int = 1):
        self.n = int(n)
        self._lock = threading.Lock()
        self._local = 0
        self._store = None
        self._buf: List[int] = []
        self._seen = 0          # highest position this rank has seen handed out (estimate of the queue head)
        self._drawers = max(1, int(drawers))
        dist = _dist()
        if dist is not None and dist.get_world_size(group) > 1:
            from torch.distributed import distributed_c10d as c10d

            self._store = c10d._get_default_store()
            self._key = f"legolas_b200/queue/{next(_QUEUE_SEQ)}"
            self._drawers *= dist.get_world_size(group)
        else:
            next(_QUEUE_SEQ)

    def next(self) -> Optional[int]:
        with self._lock:
            if not self._buf:
                if self._seen >= self.n:
                    return None
                chunk = max(1, (self.n - self._seen) // (4 * self._drawers))
                if self._store is not None:
                    end = int(self._store.add(self._key, chunk))
                else:
                    self._local += chunk
                    end = self._local
                self._seen = end
                self._buf = [p for p in range(end - chunk, end) if p < self.n]
                if not self._buf:
                    return None
            return self._buf.pop(0)

    def __iter__(self):
        while True:
            pos = self.next()
            if pos is None:
                return
            yield pos


def merge_tables(table: np.ndarray, group=None) -> np.ndarray:
    """Every rank holds an (n_units x nev) complex table with NaN rows for the units it did not solve;
    returns the union on every rank (one all_gather; nccl: device tensors, gloo: host tensors)."""
    dist = _dist()
    if dist is None or dist.get_world_size(group) == 1:
        return table
    import torch

    use_cuda = dist.get_backend(group) == "nccl"
    device = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    flat = np.ascontiguousarray(table, dtype=np.complex128).view(np.float64).reshape(-1)
    payload = torch.from_numpy(flat.copy()).to(device)
    out = [torch.empty_like(payload) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, payload, group=group)
    merged = table.copy()
    for chunk in out:
        other = chunk.cpu().numpy().view(np.complex128).reshape(table.shape)
        take = np.isnan(merged.real) & ~np.isnan(other.real)
        merged[take] = other[take]
    return merged


def gather_eigenvalues(local: np.ndarray, local_ids: Sequence[int], n_units: int, nev: int,
                       group=None) -> np.ndarray:
    """All ranks contribute ``local`` (len(local_ids) x nev complex, NaN-padded) and every rank
    gets the full (n_units x nev) table."""
    table = np.full((n_units, nev), np.nan + 1j * np.nan, dtype=np.complex128)
    local = np.asarray(local, dtype=np.complex128).reshape(len(local_ids), nev)
    for row, uid in zip(local, local_ids):
        table[uid] = row
    return merge_tables(table, group=group)


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


def run_queue(units: Sequence, make_solver: Callable[[int], Callable], nev: int, order: Optional[Sequence[int]] = None,
              workers: int = 1, group=None, gather: bool = True, solvers: Optional[list] = None):
    """Dynamic scheduling: every worker thread of every rank draws the next position of ``order`` (default: the list
    order; pass the units sorted by decreasing cost estimate for longest-first) from a shared WorkQueue and solves
    ``units[order[pos]]`` with its own solver ``make_solver(worker_index)(unit) -> omega``.

    Returns ``(table, mine)``: the (n_units x nev) table (merged over ranks if ``gather``) and the unit indices this
    rank solved.  ``solvers`` lets the caller reuse solver objects (contexts) across calls."""
    order = list(range(len(units))) if order is None else list(order)
    if sorted(order) != list(range(len(units))):
        raise ValueError("order must be a permutation of the unit indices")
    queue = WorkQueue(len(order), group=group, drawers=len(solvers) if solvers is not None else max(1, workers))
    table = np.full((len(units), nev), np.nan + 1j * np.nan, dtype=np.complex128)
    mine: List[int] = []
    errors: List[BaseException] = []
    if solvers is None:
        solvers = [make_solver(w) for w in range(max(1, workers))]

    def work(solve):
        try:
            for pos in queue:
                uid = order[pos]
                omega = np.asarray(solve(units[uid]), dtype=np.complex128)
                table[uid, :min(nev, len(omega))] = omega[:nev]
                mine.append(uid)   # list.append is atomic
        except BaseException as exc:   # noqa: BLE001 - re-raised below, in the caller's thread
            errors.append(exc)

    if len(solvers) == 1:
        work(solvers[0])
    else:
        threads = [threading.Thread(target=work, args=(s,), daemon=True) for s in solvers]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
    if gather:
        table = merge_tables(table, group=group)
    return table, sorted(mine)
