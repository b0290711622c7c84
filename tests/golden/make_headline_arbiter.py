"""Generates tests/golden/headline_arbiter.npz: eigenvalues of BASELINE config 4 at the headline
size (magnetothermal_instabilities, G = 10 001, sigma = 0.02+0.03i, nev = 20, reference defaults)
from (i) the reference-equivalent CPU path ``oracle.solvers.shift_invert`` (LAPACK zgbtrf/zgbtrs/
zgbmv + ARPACK) and (ii) the extended-precision arbiter ``oracle.solvers.shift_invert_extended``
with its default LAPACK preconditioner (every OP*x refined with 80-bit residuals until the
correction stagnates).  CPU only (about 15 minutes on 8 cores); the GPU parity test
tests/test_gpu_headline.py re-runs (i) on the GPU box and checks (ii) against a device-
preconditioned arbiter run, so this file is a cross-check, not the only pin.

    python tests/golden/make_headline_arbiter.py [gridpts]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import assembly as asm, equilibria as oeq, solvers as osolvers  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
SIGMA, NEV = 0.02 + 0.03j, 20
so, go, xgo, fo = oeq.magnetothermal_eq(gridpts=G)
A, B = asm.build_matrices(so, go, xgo, fo)
t = time.perf_counter()
om_o, vr_o, st_o = osolvers.shift_invert(A.to_band(), B.to_band(), 31, 31, SIGMA, NEV, return_stats=True)
t_o = time.perf_counter() - t
print("oracle", t_o, {k: st_o[k] for k in ("nconv", "n_op")}, flush=True)
t = time.perf_counter()
om_a, vr_a, st_a = osolvers.shift_invert_extended(A, B, SIGMA, NEV, return_stats=True)
t_a = time.perf_counter() - t
print("arbiter", t_a, st_a, flush=True)
out = os.path.join(ROOT, "tests", "golden", "headline_arbiter.npz" if G == 10001 else f"headline_arbiter_G{G}.npz")
np.savez_compressed(out, gridpts=G, sigma=SIGMA, nev=NEV, omega_oracle=om_o, omega_arbiter=om_a,
                    n_op_oracle=st_o["n_op"], nconv_oracle=st_o["nconv"], n_op_arbiter=st_a["n_op"],
                    nconv_arbiter=st_a["nconv"], sweeps_max=st_a["sweeps_max"],
                    last_correction_max=st_a["last_correction_max"], seconds=np.array([t_o, t_a]))
for w in om_a[np.argsort(abs(om_a - SIGMA))]:
    k = np.argmin(abs(om_o - w))
    print(f"{w:.14f}  oracle rel dev {abs(om_o[k]-w)/abs(w):.1e}")
