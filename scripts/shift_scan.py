"""Scratch: cost and convergence of each shift of bench.py's 8-shift scan on one GPU."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import legolas_b200 as lb
from legolas_b200 import equilibria as heq
import bench
s, grid, fields = heq.magnetothermal_instabilities(bench.GRIDPTS)
ctx = lb.Context()
cands = list(bench.SHIFTS) + [complex(x) for x in sys.argv[1:]]
for sigma in cands:
    s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=bench.NEV, sigma=sigma)
    mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
    t = time.perf_counter(); omega, vr, cfg, st = lb.solve_evp(mats, s, vr_view=True); t = time.perf_counter() - t
    print(f"sigma {sigma}: {1e3*t:8.1f} ms nconv {st['nconv']} n_op {st['n_op']} restarts {st['n_restart']} info {st['info']}")
