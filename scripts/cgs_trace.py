"""Scratch: timeline of the fused Gram-Schmidt step (library built with LGPU_NVCC_EXTRA=-DLGPU_TRACE, LGPU_LIB=...).
Per phase: min / median / max over the CTAs of the stamp, ns relative to the first CTA's start; the timeline read back
is that of the LAST step of a short Arnoldi run (basis of ncv - 1 columns)."""
import ctypes, sys
import numpy as np
sys.path.insert(0, ".")
import legolas_b200 as lb
from legolas_b200 import equilibria as heq, _lib
G = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
ncv_cols = int(sys.argv[2]) if len(sys.argv) > 2 else 0
s, grid, fields = heq.magnetothermal_instabilities(G)
s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=20, sigma=0.02 + 0.03j)
ctx = lb.Context()
mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
cfg = lb.new_arpack_config(ctx.dim, 2, "I", s.solvers); cfg.maxiter = 2
if ncv_cols: cfg.ncv = ncv_cols
ctx.shift_invert(cfg, 0.02 + 0.03j, want_vectors=False)
lib = _lib.load()
buf = np.zeros(160 * 16, dtype=np.uint64)
rc = lib.lgpu_debug_cgs_trace(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
t = buf.reshape(160, 16)
live = t[:, 0] > 0
t = t[live]
t0 = t[:, 0].min()
names = ["start", "first tile in registers", "pass 1 done", "dots published", "barrier 1 passed", "partials summed",
         "pass 2 done", "dots + norm published", "barrier 2 passed", "partials summed, norm", "pass 3 done"]
print(f"G = {G}, ncv = {cfg.ncv}, CTAs {int(live.sum())}")
prev = None
for k, nm in enumerate(names):
    col = t[:, k].astype(np.int64) - np.int64(t0)
    line = f"{k:2d} {nm:28s} min {col.min():7d} med {int(np.median(col)):7d} max {col.max():7d}"
    if prev is not None:
        line += f"   (+{int(np.median(col)) - prev} med)"
    prev = int(np.median(col))
    print(line)
