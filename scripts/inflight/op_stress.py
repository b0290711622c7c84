"""Scratch: three capped contexts in flight, each applying ITS operator to ITS fixed vector again and again - does the
output of the solve kernels ever change when neighbours run on the GPU?  (argv: repetitions per thread, mode op|solve)"""
import sys, threading
import numpy as np
sys.path.insert(0, ".")  # run from the repository root
from legolas_b200 import api, equilibria, workloads as wl
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
mode = sys.argv[2] if len(sys.argv) > 2 else "op"
units = wl.sweep_units(12)[:3]
ctxs, xs, refs = [], [], []
rng = np.random.default_rng(7)
for u in units:
    s, grid, fields = equilibria.kelvin_helmholtz_cd(wl.SWEEP_GRIDPTS, k2=u["k2"], k3=u["k3"])
    ctx = api.Context()
    ctx.set_sm_limit(148 // 3)
    api.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
    ctx.factorize(u["sigma"])
    x = rng.standard_normal(ctx.dim) + 1j * rng.standard_normal(ctx.dim)
    f = ctx.apply_op if mode == "op" else ctx.solve
    y0 = f(x)
    assert np.array_equal(y0, f(x)), "not reproducible even alone"
    ctxs.append(ctx); xs.append(x); refs.append(y0)
bad = [0, 0, 0]
worst = [0.0, 0.0, 0.0]
def work(k):
    f = ctxs[k].apply_op if mode == "op" else ctxs[k].solve
    for _ in range(reps):
        y = f(xs[k])
        if not np.array_equal(y, refs[k]):
            bad[k] += 1
            d = np.abs(y - refs[k])
            worst[k] = max(worst[k], float(d.max() / np.abs(refs[k]).max()))
            if bad[k] <= 3:
                idx = np.flatnonzero(d > 0)
                print(f"ctx {k}: {len(idx)} entries differ, first {idx[:6]}, last {idx[-3:]}, max rel {worst[k]:.3e}", flush=True)
ths = [threading.Thread(target=work, args=(k,)) for k in range(3)]
[t.start() for t in ths]; [t.join() for t in ths]
print(mode, "mismatching applications per context:", bad, "of", reps, "worst rel", worst)
