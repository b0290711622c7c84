"""Scratch: convergence and cost of candidate shifts for the multi-shift scan of config 4 (one GPU)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import legolas_b200 as lb
from legolas_b200 import equilibria as heq
G = 10001
s, grid, fields = heq.magnetothermal_instabilities(G)
ctx = lb.Context()
res = [0.0, 0.006, 0.012, 0.018, 0.024]
ims = [0.010, 0.016, 0.022, 0.028, 0.034, 0.040, 0.046]
cands = [complex(a, b) for b in ims for a in res] + [complex(-a, b) for b in (0.016, 0.028, 0.040) for a in (0.012, 0.024)]
mats = None
for sigma in cands:
    for nev in (20, 10):
        s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=nev, sigma=sigma)
        if mats is None:
            mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
        t = time.perf_counter(); omega, vr, cfg, st = lb.solve_evp(mats, s, vr_view=True); t = time.perf_counter() - t
        far = np.nanmax(np.abs(omega - sigma)) if st["nconv"] else float("nan")
        print(f"sigma {sigma.real:+.3f}{sigma.imag:+.3f}i nev {nev}: {1e3*t:8.1f} ms nconv {st['nconv']:2d} n_op {st['n_op']:5d} restarts {st['n_restart']:3d} info {st['info']} radius {far:.4f}", flush=True)
        if st["nconv"] == nev:
            break
