import sys; sys.path.insert(0, ".")
import numpy as np
import legolas_b200 as lb
from legolas_b200 import equilibria as heq
from oracle import assembly as asm, equilibria as oeq, solvers as osolvers
name, G, nev = "suydam_cluster", 1001, 10
for sigma in (-0.13 + 0.005j, -0.12 + 0.01j, -0.1 + 0.02j):
    s, grid, fields = heq.EQUILIBRIA[name](G)
    so, go, xgo, fo = oeq.EQUILIBRIA[name](gridpts=G)
    A, B = asm.build_matrices(so, go, xgo, fo)
    s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=nev, sigma=sigma)
    mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields)
    omega, vr, cfg, st = lb.solve_evp(mats, s)
    om_o, vr_o, st_o = osolvers.shift_invert(A.to_band(), B.to_band(), 31, 31, sigma, nev, return_stats=True)
    print(sigma, "gpu nconv", st["nconv"], "n_op", st["n_op"], "info", st["info"], "| oracle nconv", st_o["nconv"], "n_op", st_o["n_op"])
    print("  gpu   ", np.round(omega[np.isfinite(omega)][:4], 6))
    print("  oracle", np.round(om_o[:4], 6))
