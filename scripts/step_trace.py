"""Scratch: one Arnoldi step on a common clock (library built with -DLGPU_TRACE): the three solve kernels of the LAST
operator application of a short run and the Gram-Schmidt step that follows it, %globaltimer stamps (ns) relative to
the first CTA start of the forward first-stage kernel."""
import ctypes, sys
import numpy as np
sys.path.insert(0, ".")
import legolas_b200 as lb
from legolas_b200 import equilibria as heq, _lib
G = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
s, grid, fields = heq.magnetothermal_instabilities(G)
s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=20, sigma=0.02 + 0.03j)
ctx = lb.Context()
mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
cfg = lb.new_arpack_config(ctx.dim, 2, "I", s.solvers); cfg.maxiter = 2
ctx.shift_invert(cfg, 0.02 + 0.03j, want_vectors=False)
lib = _lib.load()
T = []
for w in range(3):
    buf = np.zeros(512 * 16, dtype=np.uint64)
    assert lib.lgpu_debug_solve_trace(ctypes.c_int(w), buf.ctypes.data_as(ctypes.c_void_p)) == 0
    T.append(buf.reshape(512, 16))
cb = np.zeros(160 * 16, dtype=np.uint64)
assert lib.lgpu_debug_cgs_trace(cb.ctypes.data_as(ctypes.c_void_p)) == 0
C = cb.reshape(160, 16); C = C[C[:, 0] > 0]
t0 = T[0][:, 0][T[0][:, 0] > 0].min()
rel = lambda a: a.astype(np.int64) - np.int64(t0)
def span(name, a):
    a = rel(a[a > 0]); print(f"{name:34s} min {a.min():8d} med {int(np.median(a)):8d} max {a.max():8d}")
span("fwd0 start", T[0][:, 0]); span("fwd0 dependency resolved", T[0][:, 1]); span("fwd0 end", T[0][:, 2])
span("upper start", T[1][:, 0]); span("upper dependency resolved", T[1][:, 2])
up_end = np.where(T[1][:, 6:10] > 0, T[1][:, 6:10], 0).max(axis=1); span("upper end", up_end)
span("bwd0 start", T[2][:, 0]); span("bwd0 dependency resolved", T[2][:, 1]); span("bwd0 end", T[2][:, 2])
names = ["start", "first tile in registers", "pass 1 done", "dots published", "barrier 1 passed", "partials summed",
         "pass 2 done", "dots + norm published", "barrier 2 passed", "partials summed, norm", "pass 3 done"]
for k, nm in enumerate(names):
    span("cgs2 " + nm, C[:, k])
