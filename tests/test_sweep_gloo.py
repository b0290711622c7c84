"""N > 1 host logic on CPU: world_size-2 gloo run of the shard / gather plumbing."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from legolas_b200 import sweep


def test_partition_round_robin():
    assert sweep.partition(8, 0, 1) == list(range(8))
    parts = [sweep.partition(256, r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == list(range(256))
    assert all(len(p) == 32 for p in parts)
    assert sweep.partition(5, 1, 2) == [1, 3]
    with pytest.raises(ValueError):
        sweep.partition(4, 2, 2)


def fake_solve(unit):
    sigma = unit
    return np.array([sigma + 1.0 / (k + 1) for k in range(3)])


def _worker(rank, world, port, units, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    table = sweep.run_sweep(units, fake_solve, nev=4, rank=rank, world_size=world)
    np.save(os.path.join(out_dir, f"table_{rank}.npy"), table)
    dist.barrier()
    dist.destroy_process_group()


def test_sweep_world2_gloo(tmp_path):
    units = [0.02 + 0.03j, 0.02 + 0.045j, 0.018 + 0.024j, 0.015 + 0.018j, 0.028j]
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    mp.spawn(_worker, args=(2, port, units, str(tmp_path)), nprocs=2, join=True)
    t0 = np.load(tmp_path / "table_0.npy")
    t1 = np.load(tmp_path / "table_1.npy")
    assert np.array_equal(t0, t1, equal_nan=True)
    single = sweep.run_sweep(units, fake_solve, nev=4)
    assert np.array_equal(t0, single, equal_nan=True)
    for i, u in enumerate(units):
        assert np.allclose(t0[i, :3], fake_solve(u))
        assert np.isnan(t0[i, 3])   # nconv < nev slots stay NaN


# ---- shared work queue (dynamic scheduling, several units in flight per rank)
def slow_solve_factory(worker):
    import time

    def solve(unit):
        time.sleep(0.002 * (1 + (int(round(unit.real * 1000)) % 5)))   # unequal costs
        return fake_solve(unit)
    return solve


def _queue_worker(rank, world, port, units, order, workers, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    for rep in range(2):   # two queues in a row: the store key is numbered per process
        table, mine = sweep.run_queue(units, slow_solve_factory, nev=4, order=order, workers=workers)
        np.save(os.path.join(out_dir, f"qtable_{rep}_{rank}.npy"), table)
        np.save(os.path.join(out_dir, f"qmine_{rep}_{rank}.npy"), np.array(mine, dtype=np.int64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("workers", [1, 3])
def test_queue_world2_gloo(tmp_path, workers):
    units = [complex(0.001 * i, 0.03) for i in range(23)]
    order = list(np.argsort([-(i % 5) for i in range(23)], kind="stable"))   # "longest first"
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    mp.spawn(_queue_worker, args=(2, port, units, order, workers, str(tmp_path)), nprocs=2, join=True)
    for rep in range(2):
        t0 = np.load(tmp_path / f"qtable_{rep}_0.npy")
        t1 = np.load(tmp_path / f"qtable_{rep}_1.npy")
        assert np.array_equal(t0, t1, equal_nan=True)
        m0 = np.load(tmp_path / f"qmine_{rep}_0.npy")
        m1 = np.load(tmp_path / f"qmine_{rep}_1.npy")
        assert sorted(np.concatenate([m0, m1]).tolist()) == list(range(23))   # every unit exactly once
        assert len(m0) > 0 and len(m1) > 0
        for i, u in enumerate(units):
            assert np.allclose(t0[i, :3], fake_solve(u)) and np.isnan(t0[i, 3])


def test_queue_single_process_matches_static_sweep():
    units = [complex(0.001 * i, 0.03) for i in range(7)]
    table, mine = sweep.run_queue(units, lambda w: fake_solve, nev=4, workers=2)
    assert mine == list(range(7))
    assert np.array_equal(table, sweep.run_sweep(units, fake_solve, nev=4), equal_nan=True)
    with pytest.raises(ValueError):
        sweep.run_queue(units, lambda w: fake_solve, nev=4, order=[0, 0, 1, 2, 3, 4, 5])


def test_queue_propagates_worker_errors():
    def factory(w):
        def solve(unit):
            raise RuntimeError("unit failed")
        return solve
    with pytest.raises(RuntimeError):
        sweep.run_queue([1.0, 2.0], factory, nev=1, workers=2)
