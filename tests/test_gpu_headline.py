"""The configuration the headline metric is quoted on (BASELINE config 4: magnetothermal_instabilities,
G = 10 001, N = 160 016, sigma = 0.02+0.03i, nev = 20, reference defaults) against the oracle, mode by mode.

Three runs of the same znaupd / zneupd call pattern (smod_arpack_shift_invert.f08:63-157), same start vector:
  device   legolas_b200 (own IRAM on the device operator)
  oracle   oracle.solvers.shift_invert: LAPACK zgbtrf / zgbtrs / zgbmv + SciPy's ARPACK (the reference-equivalent path)
  arbiter  oracle.solvers.shift_invert_extended: SciPy's ARPACK on an operator made forward-accurate by iterative
           refinement with 80-bit residuals (corrections proposed by the device solve; the result is certified by the
           extended-precision residual, not by the proposer) - cross-checked against the committed CPU-only run of the
           same arbiter with LAPACK-proposed corrections (tests/golden/headline_arbiter.npz, make_headline_arbiter.py).

Rule (the one tests/test_gpu_parity.py uses at G <= 5001): every device eigenvalue is within 1e-8 of the oracle's, or at
least as close to the arbiter as the oracle's is; same for eigenvectors at 1e-6 after phase normalisation.  Measured
(profiles/headline_parity_r2.md): at this size the LAPACK-based path is 4e-5 ... 2e-1 away from the arbiter (its band LU
loses 2 digits of forward accuracy per solve, so ARPACK also needs 343-422 operator applications instead of 188), the
device path 2e-8 ... 2e-6.
"""
import json
import os

import numpy as np
import pytest

import legolas_b200 as lb
from legolas_b200 import equilibria as heq
from oracle import assembly as asm
from oracle import equilibria as oeq
from oracle import solvers as osolvers

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, NEV, SIGMA = 10001, 20, 0.02 + 0.03j


def _phase_normalised(v):
    k = int(np.argmax(np.abs(v)))
    return v * (np.conj(v[k]) / abs(v[k])) / np.linalg.norm(v)


@pytest.fixture(scope="module")
def runs():
    s, grid, fields = heq.magnetothermal_instabilities(G)
    s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=NEV, sigma=SIGMA)
    ctx = lb.Context()
    mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
    om_g, vr_g, cfg, st_g = lb.solve_evp(mats, s)
    vr_g = np.array(vr_g)
    so, go, xgo, fo = oeq.magnetothermal_eq(gridpts=G)
    A, B = asm.build_matrices(so, go, xgo, fo)
    om_o, vr_o, st_o = osolvers.shift_invert(A.to_band(), B.to_band(), 31, 31, SIGMA, NEV, return_stats=True)
    ctx.factorize(SIGMA)
    om_a, vr_a, st_a = osolvers.shift_invert_extended(A, B, SIGMA, NEV, solve=ctx.solve, return_stats=True)
    ctx.close()
    rows = []
    for j in np.argsort(np.abs(om_a - SIGMA)):
        w = om_a[j]
        kg, ko = int(np.nanargmin(np.abs(om_g - w))), int(np.argmin(np.abs(om_o - w)))
        va = _phase_normalised(vr_a[:, j])
        rows.append({"omega": w, "kg": kg, "ko": ko,
                     "dev_gpu": abs(om_g[kg] - w) / abs(w), "dev_oracle": abs(om_o[ko] - w) / abs(w),
                     "oracle_vs_gpu": abs(om_g[kg] - om_o[ko]) / abs(w),
                     "vec_gpu": float(np.linalg.norm(_phase_normalised(vr_g[:, kg]) - va)),
                     "vec_oracle": float(np.linalg.norm(_phase_normalised(vr_o[:, ko]) - va)),
                     "vec_oracle_vs_gpu": float(np.linalg.norm(_phase_normalised(vr_g[:, kg]) - _phase_normalised(vr_o[:, ko])))})
    out = os.path.join(ROOT, "gpurun_out")
    try:   # the per-mode table, for profiles/
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "headline_modes.json"), "w") as fh:
            json.dump({"gpu": st_g, "oracle": {k: st_o[k] for k in ("nconv", "n_op")},
                       "arbiter": {k: (float(v) if isinstance(v, float) else int(v)) for k, v in st_a.items()},
                       "modes": [{k: ([v.real, v.imag] if isinstance(v, complex) else v) for k, v in r.items()} for r in rows]},
                      fh, indent=1)
    except OSError:
        pass
    return {"st_g": st_g, "st_o": st_o, "st_a": st_a, "om_g": om_g, "om_o": om_o, "om_a": om_a, "rows": rows}


def test_headline_converges_like_arpack_on_an_accurate_operator(runs):
    st_g, st_o, st_a = runs["st_g"], runs["st_o"], runs["st_a"]
    assert st_g["info"] == 0
    assert st_g["nconv"] == st_o["nconv"] == st_a["nconv"] == NEV
    # the device IRAM restates ARPACK: on the same (accurate) operator the two take the same number of operator
    # applications; the LAPACK-based path needs more because its solves are two digits less accurate
    assert abs(st_g["n_op"] - st_a["n_op"]) <= max(2, st_a["n_op"] // 50), (st_g["n_op"], st_a["n_op"])
    assert st_o["n_op"] >= st_g["n_op"]


def test_headline_modes_are_the_same_set(runs):
    rows = runs["rows"]
    assert len({r["kg"] for r in rows}) == NEV      # one device mode per arbiter mode
    # every mode of the three runs is identified unambiguously wherever the LAPACK path is still resolving the
    # sequence (its outermost members are off by 4e-2 ... 2e-1, i.e. by about the mode spacing)
    assert sum(r["dev_oracle"] <= 1e-3 for r in rows) >= 12


def test_headline_eigenvalues_against_oracle_and_arbiter(runs):
    for i, r in enumerate(runs["rows"]):
        ok = r["oracle_vs_gpu"] <= 1e-8 or r["dev_gpu"] <= r["dev_oracle"]
        assert ok, (i, r)
        assert r["dev_gpu"] <= 2e-5, (i, r)     # measured 2e-8 ... 2e-6


def test_headline_eigenvectors_against_oracle_and_arbiter(runs):
    for i, r in enumerate(runs["rows"]):
        ok = r["vec_oracle_vs_gpu"] <= 1e-6 or r["vec_gpu"] <= r["vec_oracle"]
        assert ok, (i, r)
        assert r["vec_gpu"] <= 2e-3, (i, r)     # measured 4e-9 ... 1e-4


def test_headline_arbiter_agrees_with_the_cpu_only_arbiter(runs):
    """The arbiter run above takes its correction proposals from the device solve; the committed fixture is the same
    arbiter with LAPACK proposals, computed without any GPU.  They must give the same eigenvalues."""
    fx = np.load(os.path.join(ROOT, "tests", "golden", "headline_arbiter.npz"))
    om_f = fx["omega_arbiter"]
    assert int(fx["nconv_arbiter"]) == NEV
    for w in runs["om_a"]:
        assert np.abs(om_f - w).min() <= 2e-8 * abs(w), w
