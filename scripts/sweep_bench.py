"""BASELINE config 5: kelvin_helmholtz_cd, G = 2001, 256-point (k2, k3) wavenumber sweep sharded over the
GPUs of one node (legolas_b200.sweep: round-robin units, no data-path collective, one all_gather of the
eigenvalue table).  Not the bench contract (bench.py measures the headline config); run as
    python scripts/sweep_bench.py [--check 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/sweep_bench.py
k2 in {-8 .. 7} (integers, as cylindrical geometry requires) x 16 values of k3 in [pi/4, 4 pi]; shift
sigma = sigma0 k3 / pi at the edge of the flow continuum.  At G = 2001 the
continuum is resolved so finely that only the one or two discrete Kelvin-Helmholtz modes next to the shift
converge under the reference's tolerance (5e-15) - nev = 6 runs stop at maxiter with 0 ... 4 pairs on the
device and in the CPU oracle alike - so the sweep tracks nev = 2 modes with ncv = 16 and a bounded maxiter;
units without an unstable mode end at maxiter, as they do in the reference.  --check K re-solves K units with the
CPU oracle (SciPy LAPACK + ARPACK) and reports the largest relative eigenvalue difference."""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gridpts", type=int, default=2001)
    ap.add_argument("--nev", type=int, default=2)
    ap.add_argument("--ncv", type=int, default=16)
    ap.add_argument("--maxiter", type=int, default=30)
    ap.add_argument("--sigma0", type=complex, default=2.5 + 0.8j)
    ap.add_argument("--check", type=int, default=0)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import legolas_b200 as lb
    from legolas_b200 import equilibria as heq
    from legolas_b200 import sweep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    units = [(float(k2), math.pi * 0.25 * (j + 1)) for k2 in range(-8, 8) for j in range(16)]
    ctx = lb.Context(device=local_rank)
    stats = {"nconv_min": args.nev, "n_op": 0}

    def solve_unit(unit):
        k2, k3 = unit
        s, grid, fields = heq.kelvin_helmholtz_cd(args.gridpts, k2=k2, k3=k3)
        s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=args.nev,
                                      sigma=args.sigma0 * k3 / math.pi, ncv=args.ncv, maxiter=args.maxiter)
        mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
        omega, _, _, st = lb.solve_evp(mats, s)
        stats["nconv_min"] = min(stats["nconv_min"], st["nconv"])
        stats["n_op"] += st["n_op"]
        return omega

    solve_unit(units[rank % len(units)])          # warm-up (allocations, module load)
    stats = {"nconv_min": args.nev, "n_op": 0}
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    table = sweep.run_sweep(units, solve_unit, args.nev, rank, world)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    agg = torch.tensor([float(stats["n_op"]), float(-stats["nconv_min"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        nop = agg[:1].clone()
        dist.all_reduce(nop, op=dist.ReduceOp.SUM)
        worst = agg[1:].clone()
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        agg = torch.cat([nop, worst])
    out = {"workload": f"kelvin_helmholtz_cd G={args.gridpts}, {len(units)}-point (k2,k3) sweep, nev={args.nev}",
           "n_gpus": world, "seconds": float(dt[0]), "units_per_s": len(units) / float(dt[0]),
           "ms_per_unit_per_gpu": 1e3 * float(dt[0]) * world / len(units), "n_op_total": int(agg[0]),
           "nev": args.nev, "ncv": args.ncv, "maxiter": args.maxiter, "sigma0": [args.sigma0.real, args.sigma0.imag],
           "nconv_min": int(-agg[1]), "units_with_a_converged_mode": int(np.isfinite(table).any(axis=1).sum()), "finite_rows": int(np.isfinite(table).all(axis=1).sum())}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if rank == 0 and args.check:
        from oracle import assembly as asm
        from oracle import equilibria as oeq
        from oracle import solvers as osolvers
        worst_rel, detail = 0.0, []
        for uid in np.linspace(0, len(units) - 1, args.check).astype(int):
            k2, k3 = units[uid]
            so, go, xgo, fo = oeq.kelvin_helmholtz_cd_eq(gridpts=args.gridpts, k2=k2, k3=k3)
            A, B = asm.build_matrices(so, go, xgo, fo)
            om_o, _, st_o = osolvers.shift_invert(A.to_band(), B.to_band(), 31, 31, args.sigma0 * k3 / math.pi,
                                                  args.nev, ncv=args.ncv, maxiter=args.maxiter, return_stats=True)
            got = table[uid][np.isfinite(table[uid])]
            detail.append({"unit": [k2, k3], "gpu_nconv": int(got.size), "oracle_nconv": int(st_o["nconv"])})
            for w in got:
                if len(om_o):
                    worst_rel = max(worst_rel, float(np.min(np.abs(om_o - w)) / abs(w)))
        out["oracle_check"] = {"units": detail, "max_rel_eig_diff": worst_rel}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
