"""Scratch: headline solve repeated; n_op, hashes of the eigenvalues / vectors and the iteration time of every repetition."""
import sys, hashlib
import numpy as np
sys.path.insert(0, ".")  # run from the repository root
import legolas_b200 as lb
from legolas_b200 import equilibria as heq
G = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
s, grid, fields = heq.magnetothermal_instabilities(G)
s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=20, sigma=0.02 + 0.03j)
ctx = lb.Context()
mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
for rep in range(reps):
    omega, vr, cfg, st = lb.solve_evp(mats, s, vr_view=True)
    print(rep, st["n_op"], hashlib.sha256(np.ascontiguousarray(omega).tobytes()).hexdigest()[:12],
          hashlib.sha256(np.ascontiguousarray(vr).tobytes()).hexdigest()[:12], "iter_ms %.3f" % ctx.phase_times()["iter_ms"])
