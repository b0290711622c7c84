"""Scratch: headline config on GPU vs the CPU oracle at the same G."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import legolas_b200 as lb
from legolas_b200 import equilibria as heq
from oracle import assembly as asm, equilibria as oeq, solvers as osolvers
G = int(sys.argv[1]); sigma = 0.02 + 0.03j; nev = 20
s, grid, fields = heq.magnetothermal_instabilities(G)
so, go, xgo, fo = oeq.magnetothermal_eq(gridpts=G)
A, B = asm.build_matrices(so, go, xgo, fo)
def res(w, v):
    bv = B.matvec(v); return np.linalg.norm(A.matvec(v) - w * bv) / np.linalg.norm(w * bv)
ctx = lb.Context()
out = {}
for refine in (0, 1):
    s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=nev, sigma=sigma, refine_steps=refine)
    mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
    t = time.perf_counter(); omega, vr, cfg, st = lb.solve_evp(mats, s); t = time.perf_counter() - t
    print("GPU refine", refine, "time %.1f ms" % (1e3 * t), {k: st[k] for k in ("info", "nconv", "n_op", "n_restart")})
    out[refine] = (omega, vr)
t = time.perf_counter()
om_o, vr_o, st_o = osolvers.shift_invert(A.to_band(), B.to_band(), 31, 31, sigma, nev, return_stats=True)
print("CPU oracle time %.1f s" % (time.perf_counter() - t), {k: st_o[k] for k in ("nconv", "n_op", "t_factor", "t_matvec", "t_solve", "t_iter")})
order = np.argsort(np.abs(om_o - sigma))
for j in order:
    w = om_o[j]
    line = f"{w:.12f} res_o {res(w, vr_o[:, j]):.1e}"
    for refine in (0, 1):
        om, vr = out[refine]
        k = int(np.nanargmin(np.abs(om - w)))
        line += f" | r{refine}: diff {abs(om[k]-w)/abs(w):.1e} res {res(om[k], vr[:, k]):.1e}"
    print(line)
print("GPU r0 eigenvalues:", np.array2string(out[0][0][np.argsort(np.abs(out[0][0]-sigma))][:8], precision=10))
print("truth (G=501 converged) 0.02023646+0.03219117j")
