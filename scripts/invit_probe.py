"""Scratch: shifts / tolerances for which inverse iteration at G = 10 001 needs a few solves on the device AND in the oracle."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import legolas_b200 as lb
from legolas_b200 import equilibria as heq
from oracle import assembly as asm, equilibria as oeq, solvers as osolvers
G = 10001
s, grid, fields = heq.magnetothermal_instabilities(G)
ctx = lb.Context()
s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=20, sigma=0.02 + 0.03j)
mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
so, go, xgo, fo = oeq.magnetothermal_eq(gridpts=G)
A, B = asm.build_matrices(so, go, xgo, fo)
Ab, Bb = A.to_band(), B.to_band()
mode = 0.020213655952 + 0.044145501705j
for d in (1e-3 + 0j, 2e-3j, 3e-3 - 1e-3j):
    for tol in (1e-9, 1e-10, 1e-11, 1e-12):
        sig = mode + d
        ev, x, st = ctx.inverse_iteration(sig, maxiter=30, tolerance=tol)
        ev_o, x_o, info = osolvers.inverse_iteration(Ab, Bb, 31, 31, sig, maxiter=30, tol=tol, start="solve")
        print(f"d {d} tol {tol:g}: gpu solves {st['n_op']} conv {st['info'] == 0} | cpu solves {info['iterations']} conv {info['converged']} | "
              f"rel diff {abs(ev - ev_o) / abs(ev_o):.1e} | to mode {abs(ev - mode) / abs(mode):.1e}", flush=True)
