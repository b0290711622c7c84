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


# ---- row N3: ARPACK general mode (OP = B^-1 A), tests/unit_tests/mod_test_solvers_arpack_general.pf
@pytest.mark.parametrize("which,idxs", [("LM", [4, 7, 9, 10]), ("SM", [1, 2, 3, 5]), ("LR", [7, 8, 9, 10]),
                                        ("SR", [1, 2, 3, 4]), ("LI", [4, 7, 9, 10]), ("SI", [1, 2, 3, 8])])
def test_arnoldi_general_pfunit_known_answers(ctx, which, idxs):
    from test_oracle_golden import EXPECTED_10, pencil_10
    a, b = pencil_10()
    # embed in N = 16: six decoupled rows whose eigenvalue is never among the four wanted ones
    n = 16
    pad = 1.0e-3 * (1 + 1j) if which[0] == "L" else 1.0e3 * (1 + 1j)
    ap = np.zeros((n, n), dtype=complex)
    bp = np.zeros((n, n), dtype=complex)
    ap[:10, :10], bp[:10, :10] = a, b
    for i in range(10, n):
        ap[i, i], bp[i, i] = pad, 1.0
    for M, label in ((ap, "A"), (bp, "B")):
        r, c = np.nonzero(M)
        ctx.import_coo(label, n, r + 1, c + 1, M[r, c])
    sv = lb.SolverSettings(solver="arnoldi", arpack_mode="general", number_of_eigenvalues=4,
                           maxiter=500, which_eigenvalues=which)
    cfg = lb.new_arpack_config(n, 1, "I", sv)
    omega, vr, stats = ctx.arnoldi_general(cfg)
    assert stats["nconv"] == 4
    order = np.argsort(omega.real)
    omega, vr = omega[order], vr[:, order]
    assert np.abs(omega - EXPECTED_10[np.array(idxs) - 1]).max() < 1e-12
    # A v = omega B v on the original pencil
    for k in range(4):
        assert np.linalg.norm(ap @ vr[:, k] - omega[k] * (bp @ vr[:, k])) < 1e-10
    # the factors of B are not left behind as if they were those of A - sigma B
    with pytest.raises(lb.LgpuError):
        ctx.solve(np.ones(n, dtype=complex))


@pytest.mark.parametrize("name,gridpts,which,nev", [("adiabatic_homo", 31, "LM", 6),
                                                     ("kelvin_helmholtz_cd", 51, "LM", 4)])
def test_arnoldi_general_matches_oracle(ctx, name, gridpts, which, nev):
    """solve_evp(arpack_mode="general") on assembled matrices against the oracle's ARPACK run."""
    s, grid, fields = getattr(heq, name)(gridpts)
    s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="general", number_of_eigenvalues=nev,
                                  which_eigenvalues=which, maxiter=2000)
    mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
    omega, vr, cfg, stats = lb.solve_evp(mats, s)
    so, go, xgo, fo = getattr(oeq, name + "_eq")(gridpts=gridpts)
    A, B = asm.build_matrices(so, go, xgo, fo)
    om_o, _, st_o = osolvers.arnoldi_general(A.to_band(), B.to_band(), 31, 31, nev, which=which,
                                             maxiter=2000, return_stats=True)
    assert stats["nconv"] == st_o["nconv"] == nev
    for w in omega:
        assert np.min(np.abs(om_o - w)) <= 1e-8 * abs(w)
    res = ctx.residuals(omega, vr)
    assert res.max() < 1e-9


# ------------------------------------------------------------------ N2: datfile from the device
def test_datfile_of_a_device_run_reads_back_like_the_reference_file(ctx, golden, tmp_path):
    """create_datfile with every device-backed block on (eigenfunctions, residuals, matrices) for the
    reference's stored MRI run: eigenfunctions equal the stored ones, residuals agree with the
    stored ones like the oracle's do, the matrix triplets are the oracle's in the same order."""
    from legolas_b200 import datfile as ldf
    from oracle.datfile import read_datfile
    from test_datfile import mri_run
    g, hdr, s, grid, fields, io, info, _ = mri_run(golden)
    io.write_matrices = True
    so, go, xgo, fo = oeq.mri_accretion_eq(gridpts=10, nodes=g["gauss_nodes"])     # the nodes in the file's header
    so.gauss_nodes, so.gauss_weights = g["gauss_nodes"], g["gauss_weights"]
    A, B = asm.build_matrices(so, go, xgo, fo)
    ctx.assemble(s, go, xgo, fo)
    path = ldf.create_datfile(tmp_path / "mri_gpu.dat", s, go, xgo, fo, g["eigenvalues"], ctx=ctx,
                              eigenvectors=g["eigenvectors"], io=io, info=info)
    d = read_datfile(path)
    assert np.array_equal(d["ef_written_idxs"], g["ef_written_idxs"])
    for name in d["state_vector"]:
        ref = g["ef_" + name]
        assert np.all(np.abs(d["eigenfunctions"][name] - ref) <= 1e-12 * np.abs(ref).max()), name
    ok = np.isfinite(g["residuals"]) & (g["residuals"] > 0)
    ratio = d["residuals"][ok] / g["residuals"][ok]
    assert ok.sum() > 100 and ratio.min() > 0.9 and ratio.max() < 1.1, (ratio.min(), ratio.max())
    for label, M in (("matrix_A", A), ("matrix_B", B)):
        r, c, v = M.to_coo()
        assert np.array_equal(d[label][0], r) and np.array_equal(d[label][1], c), label
        ref = v if label == "matrix_A" else v.real
        assert np.all(np.abs(d[label][2] - ref) <= 1e-12 * np.abs(ref) + 1e-15 * np.abs(ref).max()), label
