"""Scratch: per-kind kernel time of one eigenproblem at G=10001 (CUDA events around every launch)."""
import sys
sys.path.insert(0, ".")
import legolas_b200 as lb
from legolas_b200 import equilibria as heq
s, grid, fields = heq.magnetothermal_instabilities(10001)
s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=20, sigma=0.02 + 0.03j)
ctx = lb.Context()
mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
lb.solve_evp(mats, s)
ctx.set_profiling(True)
lb.solve_evp(mats, s)
p = ctx.profile()
print({k: (round(v[0], 3), int(v[1]), round(1e3 * v[0] / max(v[1], 1), 1)) for k, v in p.items() if v[1] > 0})
