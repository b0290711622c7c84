"""Scratch for compute-sanitizer racecheck on the round-2 kernels: one config-5 unit on a capped context (the path of
the in-flight sweeps) and a short headline-type run (G = 1201, nev = 20: programmatic launches, completion flags)."""
import sys
sys.path.insert(0, ".")  # run from the repository root
import numpy as np
import legolas_b200 as lb
from legolas_b200 import equilibria as heq, workloads as wl
u = wl.sweep_units(12)[3]
sv = wl.SweepSolver(sm_limit=148 // 3, gridpts=int(sys.argv[1]) if len(sys.argv) > 1 else 2001)
print("sweep unit", sv(u), flush=True)
sv.close()
s, grid, fields = heq.magnetothermal_instabilities(1201)
s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=20, sigma=0.02 + 0.03j, maxiter=3)
mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields)
omega, vr, cfg, st = lb.solve_evp(mats, s)
print("headline-type", "nconv", st["nconv"], "n_op", st["n_op"], flush=True)
