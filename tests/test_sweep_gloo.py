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
