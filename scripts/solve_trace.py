"""Scratch: timeline of one operator application's solve kernels (library built with LGPU_NVCC_EXTRA=-DLGPU_TRACE).
Prints, per kernel, the spread of CTA start / dependency-resolved / end times and, for the upper-stage kernel, the
critical path stage by stage (ns, relative to the first CTA start of forward stage 0)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, ".")
import legolas_b200 as lb
from legolas_b200 import equilibria as heq, _lib
G = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
s, grid, fields = heq.magnetothermal_instabilities(G)
s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=20, sigma=0.02 + 0.03j)
ctx = lb.Context()
mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
# a short Arnoldi run: ~100 operator applications back to back (warm instruction caches, programmatic launches in
# their steady state); the timeline read back is that of the LAST application
cfg = lb.new_arpack_config(ctx.dim, 2, "I", s.solvers); cfg.maxiter = 2
ctx.shift_invert(cfg, 0.02 + 0.03j, want_vectors=False)
lib = _lib.load()
T = []
for w in range(3):
    buf = np.zeros(512 * 16, dtype=np.uint64)
    rc = lib.lgpu_debug_solve_trace(ctypes.c_int(w), buf.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, rc
    T.append(buf.reshape(512, 16))
t0 = T[0][:, 0][T[0][:, 0] > 0].min()
def rel(a): return (a.astype(np.int64) - np.int64(t0))
for w, name in ((0, "fwd0"), (1, "upper"), (2, "bwd0")):
    t = T[w]; live = t[:, 0] > 0; n = int(live.sum())
    st, dep = rel(t[live, 0]), rel(t[live, 1 if w != 1 else 2])
    last = 2 if w != 1 else np.where(t[live] [:, 6:10] > 0, t[live][:, 6:10], 0).max(axis=1)
    en = rel(t[live, 2]) if w != 1 else rel(last)
    print(f"{name}: ctas {n} start {st.min()}..{st.max()} dep-resolved {dep.min()}..{dep.max()} end {en.min()}..{en.max()}")
    if w != 1:
        sm = t[live, 15].astype(int)
        per_sm = np.bincount(sm, minlength=148)
        dur = en - dep
        print(f"   CTAs per SM: min {per_sm.min()} max {per_sm.max()} hist {np.bincount(per_sm)}; body ns: min {dur.min()} med {int(np.median(dur))} max {dur.max()}")
        for k in (1, 2, 3):
            sel = per_sm[sm] == k
            if sel.any():
                print(f"   SMs with {k} CTAs: body med {int(np.median(dur[sel]))} end max {en[sel].max()}")
t = T[1]; live = t[:, 0] > 0
stage = (t[:, 15] >> np.uint64(32)).astype(int)
names = ["start", "prologue", "dep", "inputs", "forward", "bound/top", "bwd lvl a", "bwd lvl b", "bwd lvl c"]
for sidx in sorted(set(stage[live])):
    sel = live & (stage == sidx)
    r = rel(t[sel][:, :9]).astype(float)
    r[t[sel][:, :9] == 0] = np.nan
    print(f"upper stage {sidx} ({int(sel.sum())} CTAs):")
    r = rel(t[sel][:, :15]).astype(float)
    r[t[sel][:, :15] == 0] = np.nan
    for k, nm in enumerate(names + ["-", "b0: records", "b0: rhs done", "b0: solved", "top: Linv done", "top: 64 solved"]):
        col = r[:, k]
        if np.all(np.isnan(col)): continue
        print(f"   {nm:10s} min {np.nanmin(col):9.0f} med {np.nanmedian(col):9.0f} max {np.nanmax(col):9.0f}")
