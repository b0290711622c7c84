"""Rows N1 (residuals) and N4 (inverse iteration) of the scope table through the C ABI vs the oracle."""
import numpy as np
import pytest

import legolas_b200 as lb
from legolas_b200 import equilibria as heq
from oracle import assembly as asm
from oracle import equilibria as oeq
from oracle import solvers as osolvers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = lb.Context()
    yield c
    c.close()


def phase_distance(v, ref):
    ip = np.vdot(ref, v)
    return np.linalg.norm(v * (np.conj(ip) / abs(ip)) - ref)


@pytest.mark.parametrize("name,gridpts", [("resistive_tearing", 101), ("magnetothermal_instabilities", 333),
                                          ("adiabatic_homo", 5)])
def test_residuals_match_oracle(ctx, name, gridpts):
    s, grid, fields = heq.EQUILIBRIA[name](gridpts)
    so, go, xgo, fo = oeq.EQUILIBRIA[name](gridpts=gridpts)
    A, B = asm.build_matrices(so, go, xgo, fo)
    ctx.assemble(s, grid.base_grid, grid.gaussian_grid, fields)
    rng = np.random.default_rng(4)
    n = A.n
    vr = np.asfortranarray(rng.standard_normal((n, 4)) + 1j * rng.standard_normal((n, 4)))
    omega = np.array([0.31 - 0.22j, 4e-15 - 1e-15j, -2.0 + 0.0j, 1e-3j])
    got = ctx.residuals(omega, vr)
    ref = osolvers.residuals(A.to_band(), B.to_band(), 31, 31, omega, vr)
    assert got[1] == 0.0 and ref[1] == 0.0                    # is_zero(omega) short-circuit
    assert np.all(np.abs(got - ref) <= 1e-12 * np.abs(ref))


def test_residuals_of_converged_pairs(ctx):
    """On actual eigenpairs the residual is a cancellation: same magnitude as the oracle's, and
    small."""
    name, gridpts, sigma = "resistive_tearing", 201, 0.3 - 0.2j
    s, grid, fields = heq.EQUILIBRIA[name](gridpts)
    s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=6, sigma=sigma)
    so, go, xgo, fo = oeq.EQUILIBRIA[name](gridpts=gridpts)
    A, B = asm.build_matrices(so, go, xgo, fo)
    mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
    omega, vr, _, stats = lb.solve_evp(mats, s)
    assert stats["nconv"] == 6
    got = ctx.residuals(omega, vr)
    ref = osolvers.residuals(A.to_band(), B.to_band(), 31, 31, omega, vr)
    assert np.all(got <= 1e-8)
    assert np.all(np.abs(got - ref) <= 1e-3 * ref + 1e-15)


@pytest.mark.parametrize("name,gridpts,near", [("resistive_tearing", 201, 0.3 - 0.2j),
                                               ("magnetothermal_instabilities", 333, 0.02 + 0.03j),
                                               ("kelvin_helmholtz_cd", 51, 2.5 + 0.5j)])
def test_inverse_iteration_matches_oracle(ctx, name, gridpts, near):
    s, grid, fields = heq.EQUILIBRIA[name](gridpts)
    so, go, xgo, fo = oeq.EQUILIBRIA[name](gridpts=gridpts)
    A, B = asm.build_matrices(so, go, xgo, fo)
    Ab, Bb = A.to_band(), B.to_band()
    # shift next to an eigenvalue found by the shift-invert oracle, so that both runs converge to
    # the same simple eigenvalue
    om, _ = osolvers.shift_invert(Ab, Bb, 31, 31, near, 4)
    target = om[np.argmin(np.abs(om - near))]
    sigma = complex(target * (1 + 1e-3))
    tol = 1e-10
    ev_o, x_o, info_o = osolvers.inverse_iteration(Ab, Bb, 31, 31, sigma, maxiter=60, tol=tol)
    ev_s, x_s, info_s = osolvers.inverse_iteration(Ab, Bb, 31, 31, sigma, maxiter=60, tol=tol, start="solve")
    s.solvers = lb.SolverSettings(solver="inverse-iteration", sigma=sigma, maxiter=60, tolerance=tol)
    mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
    omega, vr, _, stats = lb.solve_evp(mats, s)
    assert info_o["converged"] and info_s["converged"] and stats["info"] == 0, (info_o, info_s, stats)
    # same eigenvalue as the reference algorithm (LAPACK start vector) to the parity bar, and the
    # same iteration count as the oracle run that shares the device's start vector (+-1: the
    # stopping test sits at rounding level)
    assert abs(omega[0] - ev_o) <= 1e-8 * abs(ev_o), (omega[0], ev_o)
    assert abs(stats["n_op"] - info_s["iterations"]) <= 1, (stats, info_s)
    x = vr[:, 0]
    im = int(np.argmax(np.abs(x)))
    assert abs(x[im].imag) <= 1e-12 * abs(x[im]) and x[im].real > 0
    assert abs(np.linalg.norm(x) - 1.0) <= 1e-12
    assert phase_distance(x, x_o) <= 1e-6, phase_distance(x, x_o)
    assert ctx.residuals(omega, vr)[0] <= 1e-8


def test_inverse_iteration_argument_checks(ctx):
    s, grid, fields = heq.EQUILIBRIA["adiabatic_homo"](11)
    mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
    for bad in (dict(sigma=0j), dict(sigma=1.0 + 0j, maxiter=-3)):
        s.solvers = lb.SolverSettings(solver="inverse-iteration", **bad)
        with pytest.raises(lb.LegolasError):
            lb.solve_evp(mats, s)
    with pytest.raises(lb._lib.LgpuError):
        ctx.inverse_iteration(0j)
    # default tolerance (5e-15) is unreachable: maxiter + 1 solves, eigenvalue still returned
    omega, x, stats = ctx.inverse_iteration(1.1 + 0.2j, maxiter=5)
    assert stats["info"] == 1 and stats["n_op"] == 6 and np.isfinite(omega)


# ------------------------------------------------------------------ N1: eigenfunction assembly
def test_eigenfunctions_replay_reference_stored_run(ctx, golden):
    """The reference's own stored run (v2.0.0_mri_subset_efs.dat): its eigenvectors through the
    device assembly give the eigenfunctions it wrote."""
    g = golden("mri_subset_efs")
    s, grid, fields = heq.mri_accretion(10)
    assert np.allclose(grid.base_grid, g["grid"], rtol=0, atol=1e-14)
    ctx.assemble(s, grid.base_grid, grid.gaussian_grid, fields)
    efs = ctx.eigenfunctions(g["eigenvectors"], g["ef_written_idxs"])
    for name in (str(x) for x in g["state_vector"]):
        ref = g["ef_" + name]
        assert efs[name].shape == ref.shape
        assert np.all(np.abs(efs[name] - ref) <= 1e-12 * np.abs(ref).max()), name


@pytest.mark.parametrize("name,gridpts", [("resistive_tearing", 51), ("magnetothermal_instabilities", 1001),
                                          ("adiabatic_homo", 2)])
def test_eigenfunctions_match_oracle(ctx, name, gridpts):
    from oracle import eigenfunctions as oef
    s, grid, fields = heq.EQUILIBRIA[name](gridpts)
    ctx.assemble(s, grid.base_grid, grid.gaussian_grid, fields)
    rng = np.random.default_rng(9)
    n = 16 * gridpts
    vr = np.asfortranarray(rng.standard_normal((n, 5)) + 1j * rng.standard_normal((n, 5)))
    idxs = np.array([4, 1, 5], dtype=np.int32)          # 1-based, unordered subset
    got = ctx.eigenfunctions(vr, idxs)
    ref = oef.base_eigenfunctions(s.geometry, asm.STATE_VECTORS["mhd"], grid.base_grid, vr, idxs - 1)
    for var in asm.STATE_VECTORS["mhd"]:
        assert got[var].shape == (2 * gridpts - 1, 3)
        assert np.all(np.abs(got[var] - ref[var]) <= 1e-12 * np.abs(ref[var]).max()), var


def test_eigenfunctions_of_ritz_vectors_within_parity_bar(ctx):
    """Eigenfunctions of the device Ritz vectors vs those of the oracle's, phase-normalised: the 1e-6
    bar of the path."""
    from oracle import eigenfunctions as oef
    name, gridpts, sigma = "resistive_tearing", 201, 0.3 - 0.2j
    s, grid, fields = heq.EQUILIBRIA[name](gridpts)
    s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=4, sigma=sigma)
    so, go, xgo, fo = oeq.EQUILIBRIA[name](gridpts=gridpts)
    A, B = asm.build_matrices(so, go, xgo, fo)
    om_o, vr_o = osolvers.shift_invert(A.to_band(), B.to_band(), 31, 31, sigma, 4)
    mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
    omega, vr, _, stats = lb.solve_evp(mats, s)
    assert stats["nconv"] == 4
    got = ctx.eigenfunctions(vr, np.arange(1, 5, dtype=np.int32))
    for k in range(4):
        j = int(np.argmin(np.abs(om_o - omega[k])))
        assert abs(om_o[j] - omega[k]) <= 1e-8 * abs(omega[k])
        ip = np.vdot(vr_o[:, j], vr[:, k])
        phase = np.conj(ip) / abs(ip)
        ref = oef.base_eigenfunctions(s.geometry, asm.STATE_VECTORS["mhd"], grid.base_grid, vr_o, [j])
        for var in asm.STATE_VECTORS["mhd"]:
            scale = max(np.abs(ref[var]).max(), 1e-300)
            assert np.abs(got[var][:, k] * phase - ref[var][:, 0]).max() <= 1e-6 * max(scale, np.abs(ref["v1"]).max()), var


def test_eigenfunctions_need_an_assembled_grid(ctx):
    s, grid, fields = heq.EQUILIBRIA["adiabatic_homo"](11)
    ctx.assemble(s, grid.base_grid, grid.gaussian_grid, fields)
    r, c, v = ctx.export_coo("B")
    ctx.import_coo("B", 176, r, c, v)                    # imported matrices carry no grid
    with pytest.raises(lb.LgpuError):
        ctx.eigenfunctions(np.zeros((176, 1), dtype=np.complex128), np.array([1], dtype=np.int32))
