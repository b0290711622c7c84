"""Small end-to-end run for compute-sanitizer (memcheck / synccheck): every kernel class of the hot path once."""
import sys
sys.path.insert(0, ".")
import numpy as np
import legolas_b200 as lb
from legolas_b200 import equilibria as heq
for name, G, sigma, nev in (("magnetothermal_instabilities", 2501, 0.02 + 0.03j, 8), ("resistive_tearing", 301, 0.3 - 0.2j, 6)):
    s, grid, fields = heq.EQUILIBRIA[name](G)
    s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=nev, sigma=sigma, maxiter=4)
    mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields)
    omega, vr, cfg, st = lb.solve_evp(mats, s)
    print(name, G, "nconv", st["nconv"], "n_op", st["n_op"], "finite", bool(np.all(np.isfinite(vr))))
