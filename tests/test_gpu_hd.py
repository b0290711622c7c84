"""physics_type "hd" / "hd-1d" (row A5 of the scope table: reduced state vectors) through the C ABI.

The reference numbers 2 nb_eqs rows per grid point (10 for hd, 6 for hd-1d); every vector, index and
block crossing the ABI uses that numbering.  Oracle: oracle.assembly with the same physics_type, pinned
by the reference's stored Couette-flow HD spectrum (tests/test_oracle_golden.py)."""
import numpy as np
import pytest

import legolas_b200 as lb
from legolas_b200 import equilibria as heq
from oracle import assembly as asm
from oracle import eigenfunctions as oef
from oracle import equilibria as oeq
from oracle import solvers as osolvers
from test_gpu_parity import check_matrices, phase_distance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = lb.Context()
    yield c
    c.close()


def both(physics_type, gridpts, legacy=False, **kw):
    s, grid, fields = heq.couette_flow(gridpts, physics_type=physics_type, **kw)
    so, go, xgo, fo = oeq.couette_flow_eq(gridpts=gridpts, physics_type=physics_type, **kw)
    if legacy:
        so, go, xgo, fo = oeq.couette_flow_eq(gridpts=gridpts, physics_type=physics_type,
                                              nodes=asm.LEGACY_GAUSS_NODES, **kw)
        so.gauss_nodes, so.gauss_weights = asm.LEGACY_GAUSS_NODES, asm.LEGACY_GAUSS_WEIGHTS
        s.gauss_nodes, s.gauss_weights = so.gauss_nodes, so.gauss_weights
        grid.base_grid, grid.gaussian_grid, fields = go, xgo, fo
    return (s, grid, fields), asm.build_matrices(so, go, xgo, fo)


@pytest.mark.parametrize("physics_type,d", [("hd", 10), ("hd-1d", 6), ("mhd", 16)])
@pytest.mark.parametrize("gridpts", [2, 9, 51])
def test_assembly_matches_oracle(ctx, physics_type, d, gridpts):
    (s, grid, fields), (A, B) = both(physics_type, gridpts)
    ctx.assemble(s, grid.base_grid, grid.gaussian_grid, fields)
    assert ctx.dim == gridpts * d == A.n and ctx.dim_subblock == d
    check_matrices(ctx, A, B)


@pytest.mark.parametrize("physics_type", ["hd", "hd-1d"])
def test_matvec_solve_and_operator(ctx, physics_type):
    rng = np.random.default_rng(11)
    (s, grid, fields), (A, B) = both(physics_type, 77)
    ctx.assemble(s, grid.base_grid, grid.gaussian_grid, fields)
    n, kl = A.n, 2 * A.d - 1
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    for M, label in ((A, "A"), (B, "B")):
        y, yo = ctx.matvec(label, x), M.matvec(x)
        assert np.abs(y - yo).max() <= 1e-12 * np.abs(yo).max()
    sigma = 0.5 - 0.3j
    assert ctx.factorize(sigma) == 0
    lu = osolvers.BandedLU(A.to_band() - sigma * B.to_band(), kl, kl)
    b = B.matvec(x)
    xs, xl = ctx.solve(b), lu.solve(b)
    assert np.linalg.norm(xs - xl) <= 1e-9 * np.linalg.norm(xl)
    assert np.linalg.norm(ctx.apply_op(x) - xl) <= 1e-9 * np.linalg.norm(xl)


@pytest.mark.parametrize("physics_type,gridpts,sigma,nev", [("hd", 51, 0.53 - 0.31j, 6), ("hd", 201, 0.53 - 0.31j, 10),
                                                            ("hd-1d", 51, 10.0 + 1.0j, 3)])
def test_shift_invert_matches_oracle(ctx, physics_type, gridpts, sigma, nev):
    (s, grid, fields), (A, B) = both(physics_type, gridpts)
    s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=nev,
                                  sigma=sigma, maxiter=500)
    mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
    omega, vr, cfg, stats = lb.solve_evp(mats, s)
    kl = 2 * A.d - 1
    om_o, vr_o, st_o = osolvers.shift_invert(A.to_band(), B.to_band(), kl, kl, sigma, nev, maxiter=500,
                                             return_stats=True)
    assert cfg.evpdim == A.n and vr.shape == (A.n, nev)
    assert stats["nconv"] == st_o["nconv"] == nev
    for k, w in enumerate(omega):
        j = int(np.argmin(np.abs(om_o - w)))
        assert abs(om_o[j] - w) <= 1e-8 * abs(w)
        assert abs(np.linalg.norm(vr[:, k]) - 1.0) < 1e-12
        assert phase_distance(vr[:, k], vr_o[:, j] / np.linalg.norm(vr_o[:, j])) < 1e-6
    # residuals and eigenfunctions of these pairs, in the reduced numbering
    res = ctx.residuals(omega, vr)
    assert np.allclose(res, osolvers.residuals(A.to_band(), B.to_band(), kl, kl, omega, vr), rtol=1e-3, atol=1e-13)
    assert res.max() < 1e-9
    efs = ctx.eigenfunctions(vr, np.arange(1, nev + 1))
    efs_o = oef.base_eigenfunctions(s.geometry, ctx.state_vector, grid.base_grid, vr, range(nev))
    assert tuple(efs) == ctx.state_vector == asm.STATE_VECTORS[physics_type]
    for name in efs:
        assert np.abs(efs[name] - efs_o[name]).max() <= 1e-12 * max(np.abs(efs_o[name]).max(), 1e-300)


def test_hd_golden_spectrum_by_shift_invert(ctx, golden):
    """Eigenvalues of the reference's stored Couette-flow HD run (QR-invert, 510 values) recovered by
    device shift-invert runs around four of its listed modes (test_couette_flow_HD.py:5-13)."""
    g = golden("couette_HD_QR")
    gold = g["eigenvalues"]
    (s, grid, fields), _ = both("hd", 51, legacy=True)
    mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
    for target in (0.48448 - 0.28639j, 0.64598 - 0.20539j, 0.35402 - 0.20539j, 0.50000 - 0.44519j):
        s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=4,
                                      sigma=target + 0.003, maxiter=500)
        omega, _, _, stats = lb.solve_evp(mats, s)
        assert stats["nconv"] == 4
        assert np.min(np.abs(omega - target)) < 1e-4          # the mode quoted by the reference test
        for w in omega:                                        # and every value is one of the stored 510
            assert np.min(np.abs(gold - w)) <= 1e-8 * abs(w)


def test_hd_inverse_iteration_and_general_mode(ctx):
    (s, grid, fields), (A, B) = both("hd", 51)
    mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
    kl = 2 * A.d - 1
    om_si, _ = osolvers.shift_invert(A.to_band(), B.to_band(), kl, kl, 0.5 - 0.44j, 2, maxiter=500)
    ev, x, st = ctx.inverse_iteration(0.5 - 0.44j, maxiter=50, tolerance=1e-11)
    assert st["info"] == 0 and x.shape == (A.n,)
    assert np.min(np.abs(ev - om_si)) <= 1e-8 * abs(ev)
    assert np.linalg.norm(A.matvec(x) - ev * B.matvec(x)) <= 1e-8 * np.linalg.norm(A.matvec(x))
    s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="general", number_of_eigenvalues=4, maxiter=2000)
    omega, vr, cfg, stats = lb.solve_evp(mats, s)
    om_o, _ = osolvers.arnoldi_general(A.to_band(), B.to_band(), kl, kl, 4, maxiter=2000)
    assert stats["nconv"] == 4
    for w in omega:
        assert np.min(np.abs(om_o - w)) <= 1e-8 * abs(w)
