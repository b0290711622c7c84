"""Headline configuration (BASELINE config 4, G = 10 001, sigma = 0.02+0.03i, nev = 20) on the GPU next to
(i) the reference-equivalent CPU path (oracle.solvers.shift_invert: LAPACK zgbtrf/zgbtrs/zgbmv + ARPACK,
smod_arpack_shift_invert.f08:63-157), (ii) the extended-precision arbiter, (iii) SciPy's ARPACK driving
the device operator (explains the OP*x counts).  Writes gpurun_out/headline_parity.{json,md}.

    python scripts/headline_parity.py [gridpts]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import legolas_b200 as lb  # noqa: E402
from legolas_b200 import equilibria as heq  # noqa: E402
from oracle import assembly as asm, equilibria as oeq, solvers as osolvers  # noqa: E402
from scipy.sparse.linalg import ArpackNoConvergence, LinearOperator, eigs  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
SIGMA, NEV = 0.02 + 0.03j, 20


def phase_normalised(v):
    k = int(np.argmax(np.abs(v)))
    return v * (np.conj(v[k]) / abs(v[k])) / np.linalg.norm(v)


def main():
    out = {"gridpts": G, "sigma": [SIGMA.real, SIGMA.imag], "nev": NEV}
    s, grid, fields = heq.magnetothermal_instabilities(G)
    s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=NEV, sigma=SIGMA)
    ctx = lb.Context()
    mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
    t = time.perf_counter()
    om_g, vr_g, cfg, st_g = lb.solve_evp(mats, s)
    out["gpu"] = {"seconds": time.perf_counter() - t, **{k: st_g[k] for k in ("info", "nconv", "n_op", "n_restart", "n_reorth")}}
    print("gpu", out["gpu"], flush=True)

    so, go, xgo, fo = oeq.magnetothermal_eq(gridpts=G)
    A, B = asm.build_matrices(so, go, xgo, fo)
    Ab, Bb = A.to_band(), B.to_band()
    t = time.perf_counter()
    om_o, vr_o, st_o = osolvers.shift_invert(Ab, Bb, 31, 31, SIGMA, NEV, return_stats=True)
    out["oracle"] = {"seconds": time.perf_counter() - t, "nconv": st_o["nconv"], "n_op": st_o["n_op"]}
    print("oracle", out["oracle"], flush=True)

    # arbiter, corrections proposed by the device solve (its accuracy is certified by the 80-bit residual)
    ctx.factorize(SIGMA)
    t = time.perf_counter()
    om_a, vr_a, st_a = osolvers.shift_invert_extended(A, B, SIGMA, NEV, solve=ctx.solve, return_stats=True)
    out["arbiter_device_precond"] = {"seconds": time.perf_counter() - t, **{k: (float(v) if isinstance(v, float) else int(v)) for k, v in st_a.items()}}
    print("arbiter", out["arbiter_device_precond"], flush=True)

    # SciPy's ARPACK on the device operator: same znaupd / zneupd as the oracle, only OP differs
    n_dev = [0]

    def op(x):
        n_dev[0] += 1
        return ctx.apply_op(x)

    v0 = osolvers.zlarnv(A.n)
    t = time.perf_counter()
    try:
        nu, _ = eigs(LinearOperator((A.n, A.n), matvec=op, dtype=np.complex128), k=NEV, which="LM", ncv=2 * NEV,
                     maxiter=max(100, 10 * NEV), tol=5e-15, v0=v0.copy())
    except ArpackNoConvergence as exc:
        nu = exc.eigenvalues
    om_s = SIGMA + 1.0 / nu
    out["scipy_arpack_on_device_op"] = {"seconds": time.perf_counter() - t, "n_op": n_dev[0], "nconv": len(nu)}
    print("scipy ARPACK on device OP", out["scipy_arpack_on_device_op"], flush=True)

    fixture = os.path.join(ROOT, "tests", "golden", "headline_arbiter.npz")
    om_f = np.load(fixture)["omega_arbiter"] if (G == 10001 and os.path.exists(fixture)) else None

    rows = []
    order = np.argsort(np.abs(om_a - SIGMA))
    for j in order:
        w = om_a[j]
        kg = int(np.nanargmin(np.abs(om_g - w)))
        ko = int(np.argmin(np.abs(om_o - w)))
        ks = int(np.argmin(np.abs(om_s - w)))
        va = phase_normalised(vr_a[:, j])
        row = {"omega_arbiter": [w.real, w.imag],
               "dev_gpu": abs(om_g[kg] - w) / abs(w), "dev_oracle": abs(om_o[ko] - w) / abs(w),
               "dev_scipy_on_device_op": abs(om_s[ks] - w) / abs(w),
               "vec_gpu": float(np.linalg.norm(phase_normalised(vr_g[:, kg]) - va)),
               "vec_oracle": float(np.linalg.norm(phase_normalised(vr_o[:, ko]) - va))}
        if om_f is not None:
            row["dev_fixture_arbiter"] = float(np.abs(om_f - w).min() / abs(w))
        rows.append(row)
    out["modes"] = rows
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "headline_parity.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    with open(os.path.join(ROOT, "gpurun_out", "headline_parity.md"), "w") as fh:
        fh.write(f"# Headline parity, G = {G}, sigma = {SIGMA}, nev = {NEV}\n\n")
        for k in ("gpu", "oracle", "arbiter_device_precond", "scipy_arpack_on_device_op"):
            fh.write(f"* {k}: {out[k]}\n")
        fh.write("\n| # | omega (arbiter) | device rel. dev | oracle (LAPACK+ARPACK) rel. dev | SciPy ARPACK on device OP | "
                 "device vec | oracle vec | LAPACK-precond. arbiter (fixture) |\n|---|---|---|---|---|---|---|---|\n")
        for i, r in enumerate(rows):
            w = complex(*r["omega_arbiter"])
            fh.write(f"| {i} | {w:.12f} | {r['dev_gpu']:.1e} | {r['dev_oracle']:.1e} | {r['dev_scipy_on_device_op']:.1e} | "
                     f"{r['vec_gpu']:.1e} | {r['vec_oracle']:.1e} | {r.get('dev_fixture_arbiter', float('nan')):.1e} |\n")
    print(open(os.path.join(ROOT, "gpurun_out", "headline_parity.md")).read())


if __name__ == "__main__":
    main()
