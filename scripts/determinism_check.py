"""Scratch: the headline solve repeated in one process - are n_op and the eigenvalue bits the same every time?"""
import sys, hashlib
import numpy as np
sys.path.insert(0, ".")
import legolas_b200 as lb
from legolas_b200 import equilibria as heq
G = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sigma = 0.02 + 0.03j
s, grid, fields = heq.magnetothermal_instabilities(G)
s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=20, sigma=sigma)
ctx = lb.Context()
mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
seen = {}
for rep in range(reps):
    omega, vr, cfg, stats = lb.solve_evp(mats, s)
    key = (stats["n_op"], hashlib.sha256(np.ascontiguousarray(omega).tobytes()).hexdigest()[:12],
           hashlib.sha256(np.ascontiguousarray(vr).tobytes()).hexdigest()[:12])
    seen[key] = seen.get(key, 0) + 1
    print(rep, key)
print("distinct outcomes:", len(seen))
