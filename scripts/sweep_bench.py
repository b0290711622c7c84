"""BASELINE configs 4 and 5 as sharded workloads, outside the bench contract (bench.py carries the same two sections in
its JSON line): the 32-shift scan and the 256-point (k2, k3) sweep of legolas_b200.workloads, handed out from a shared
queue (legolas_b200.sweep.run_queue).
    python scripts/sweep_bench.py [--workers 3] [--check 6]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/sweep_bench.py
--check K re-solves K converged sweep units with the CPU oracle (SciPy LAPACK + ARPACK, same nev / ncv / maxiter) and
reports nconv and the largest relative eigenvalue difference."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=3)
    ap.add_argument("--units", type=int, default=0)
    ap.add_argument("--check", type=int, default=0)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import bench
    from legolas_b200 import workloads as wl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    ns = argparse.Namespace(sweep_units=args.units, sweep_workers=args.workers)
    out = bench.run_sharded_sections(ns, rank, world, local_rank, device)
    out["n_gpus"] = world
    if rank == 0 and args.check:
        from oracle import assembly as asm
        from oracle import equilibria as oeq
        from oracle import solvers as osolvers
        units = wl.sweep_units(args.units)
        solver = wl.SweepSolver(device=local_rank)
        got = [solver(u) for u in units]
        conv = [i for i, g in enumerate(got) if np.isfinite(g[0])]
        detail, worst = [], 0.0
        for i in [conv[j] for j in np.linspace(0, len(conv) - 1, args.check).astype(int)]:
            u = units[i]
            so, go, xgo, fo = oeq.kelvin_helmholtz_cd_eq(gridpts=wl.SWEEP_GRIDPTS, k2=u["k2"], k3=u["k3"])
            A, B = asm.build_matrices(so, go, xgo, fo)
            t0 = time.perf_counter()
            om_o, _, st_o = osolvers.shift_invert(A.to_band(), B.to_band(), 31, 31, u["sigma"], wl.SWEEP_NEV, ncv=wl.SWEEP_NCV,
                                                  maxiter=wl.SWEEP_MAXITER, return_stats=True)
            rel = float(abs(om_o[0] - got[i][0]) / abs(om_o[0])) if st_o["nconv"] else float("nan")
            worst = max(worst, rel)
            detail.append({"k2": u["k2"], "k3": u["k3"], "omega": [got[i][0].real, got[i][0].imag], "oracle_nconv": st_o["nconv"],
                           "rel_diff": rel, "oracle_s": time.perf_counter() - t0, "oracle_n_op": st_o["n_op"]})
        out["oracle_check"] = {"units": detail, "max_rel_eig_diff": worst}
        solver.close()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
