"""Scratch timing of the headline config through the public API (not the bench contract)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import legolas_b200 as lb
from legolas_b200 import equilibria as heq

G = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
sigma = complex(sys.argv[2]) if len(sys.argv) > 2 else 0.02 + 0.03j
nev = int(sys.argv[3]) if len(sys.argv) > 3 else 20
refine = int(sys.argv[4]) if len(sys.argv) > 4 else 0
s, grid, fields = heq.magnetothermal_instabilities(G)
s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=nev, sigma=sigma, refine_steps=refine)
ctx = lb.Context()
for rep in range(3):
    t0 = time.perf_counter()
    mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
    t1 = time.perf_counter()
    omega, vr, cfg, stats = lb.solve_evp(mats, s)
    t2 = time.perf_counter()
    print(f"rep {rep}: build {1e3*(t1-t0):.2f} ms solve {1e3*(t2-t1):.2f} ms", ctx.phase_times(), stats, "launches", ctx.counters(reset=True))
A = ctx.export_blocks("A"); B = ctx.export_blocks("B")
def mv(M, x):
    xb = x.reshape(-1, 16); y = np.einsum("bij,bj->bi", M[:, 1], xb)
    y[1:] += np.einsum("bij,bj->bi", M[1:, 0], xb[:-1]); y[:-1] += np.einsum("bij,bj->bi", M[:-1, 2], xb[1:])
    return y.reshape(-1)
for k in range(stats["nconv"]):
    bv = mv(B, vr[:, k]); r = mv(A, vr[:, k]) - omega[k] * bv
    print(k, omega[k], "res", np.linalg.norm(r) / np.linalg.norm(omega[k] * bv))
