"""Short workload for ncu: a few operator applications and Arnoldi steps at G = 10001."""
import sys
import numpy as np
sys.path.insert(0, ".")
import legolas_b200 as lb
from legolas_b200 import equilibria as heq
G = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
nev = int(sys.argv[2]) if len(sys.argv) > 2 else 20
maxiter = int(sys.argv[3]) if len(sys.argv) > 3 else 2
s, grid, fields = heq.magnetothermal_instabilities(G)
s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=nev,
                              sigma=0.02 + 0.03j, maxiter=maxiter)
ctx = lb.Context()
mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
cfg = lb.new_arpack_config(ctx.dim, 2, "I", s.solvers)
cfg.maxiter = maxiter
omega, vr, stats = ctx.shift_invert(cfg, 0.02 + 0.03j)
print(stats)
