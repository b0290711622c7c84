"""Scratch: one config-5 unit on the device under different kernel paths (environment switches are read at first use,
so every variant runs in its own process: python scripts/diag_sweep_unit.py [k2 j nev ncv tol])."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import legolas_b200 as lb
from legolas_b200 import equilibria as heq, workloads as wl
k2 = float(sys.argv[1]) if len(sys.argv) > 1 else -1.0
j = int(sys.argv[2]) if len(sys.argv) > 2 else 15
nev = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ncv = int(sys.argv[4]) if len(sys.argv) > 4 else 16
tol = float(sys.argv[5]) if len(sys.argv) > 5 and not sys.argv[5].startswith("-") else 5.0e-15
unit = next(u for u in wl.sweep_units() if u["k2"] == k2 and abs(u["k3"] - np.pi * (j + 1) / 16) < 1e-12)
s, grid, fields = heq.kelvin_helmholtz_cd(wl.SWEEP_GRIDPTS, k2=unit["k2"], k3=unit["k3"])
s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=nev, sigma=unit["sigma"], ncv=ncv,
                              maxiter=20, tolerance=tol)
ctx = lb.Context()
mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
import warnings
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    omega, vr, cfg, st = lb.solve_evp(mats, s)
env = {k: v for k, v in os.environ.items() if k.startswith("LGPU_")}
print(env, "nev", nev, "ncv", ncv, "tol", tol, "->", {k: st[k] for k in ("info", "nconv", "n_op", "n_restart")}, omega[:nev])
if "--op" in sys.argv:
    from oracle import assembly as asm, equilibria as oeq, solvers as osolvers
    so, go, xgo, fo = oeq.kelvin_helmholtz_cd_eq(gridpts=wl.SWEEP_GRIDPTS, k2=unit["k2"], k3=unit["k3"])
    A, B = asm.build_matrices(so, go, xgo, fo)
    Ab, Bb = A.to_band(), B.to_band()
    lu = osolvers.BandedLU(Ab - unit["sigma"] * Bb, 31, 31)
    x = osolvers.zlarnv(ctx.dim)
    ctx.factorize(unit["sigma"])
    y = ctx.apply_op(x)
    yo = lu.solve(osolvers.banded_matvec(Bb, 31, 31, x))
    print("   OP x: device vs LAPACK rel diff %.2e" % (np.linalg.norm(y - yo) / np.linalg.norm(yo)))
