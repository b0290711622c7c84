"""Host-side IRAM control flow (iram.hpp / dense_host.hpp) against SciPy's ARPACK, on CPU.

The device operations are replaced by a TEST-ONLY C++ double (tests/csrc/iram_cpu.cpp);
the product library contains no such path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg
from scipy.sparse.linalg import LinearOperator, eigs

from oracle import assembly as asm
from oracle import equilibria as oeq
from oracle import solvers as osolvers

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "iram_cpu.cpp")
OUT = os.path.join(HERE, "_build", "libiram_cpu.so")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC] + [os.path.join(HERE, "..", "legolas_b200", "csrc", f) for f in ("iram.hpp", "dense_host.hpp")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", SRC, "-o", OUT])
    return C.CDLL(OUT)


OPFN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)


def run_iram(lib, n, matvec, v0, nev, ncv, maxiter, which="LM", tol=5e-15):
    count = [0]

    def cb(xp, yp):
        x = np.ctypeslib.as_array(C.cast(xp, C.POINTER(C.c_double)), shape=(2 * n,)).view(np.complex128)
        y = np.ctypeslib.as_array(C.cast(yp, C.POINTER(C.c_double)), shape=(2 * n,)).view(np.complex128)
        y[:] = matvec(x.copy())
        count[0] += 1

    fn = OPFN(cb)
    ritz = np.zeros(nev, dtype=np.complex128)
    vecs = np.zeros((n, nev), dtype=np.complex128, order="F")
    stats = (C.c_int * 4)()
    v0 = np.ascontiguousarray(v0, dtype=np.complex128)
    lib.iram_cpu_run(n, fn, v0.ctypes.data_as(C.c_void_p), nev, ncv, maxiter, which.encode(),
                     C.c_double(tol), ritz.ctypes.data_as(C.c_void_p),
                     vecs.ctypes.data_as(C.c_void_p), stats)
    info, nconv, n_op, n_iter = list(stats)
    assert n_op == count[0]
    return ritz[:nconv], vecs[:, :nconv], dict(info=info, nconv=nconv, n_op=n_op, n_iter=n_iter)


def test_hessenberg_schur_and_eigvecs(lib):
    rng = np.random.default_rng(0)
    for n in (1, 2, 5, 12, 40):
        H = np.triu(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)), -1)
        T = np.asfortranarray(H.copy())
        Z = np.zeros((n, n), dtype=complex, order="F")
        w = np.zeros(n, dtype=complex)
        rc = lib.dense_hessenberg_schur(n, T.ctypes.data_as(C.c_void_p), Z.ctypes.data_as(C.c_void_p),
                                        w.ctypes.data_as(C.c_void_p))
        assert rc == 0
        assert np.abs(np.tril(T, -1)).max() == 0 if n > 1 else True
        assert np.allclose(Z.conj().T @ Z, np.eye(n), atol=1e-13)
        assert np.allclose(Z @ T @ Z.conj().T, H, atol=1e-12 * max(1, np.abs(H).max()))
        ref = np.linalg.eigvals(H)
        assert max(np.min(np.abs(ref - x)) for x in w) < 1e-11
        assert np.allclose(np.diag(T), w)
        X = np.zeros((n, n), dtype=complex, order="F")
        lib.dense_triangular_eigvecs(n, T.ctypes.data_as(C.c_void_p), n, X.ctypes.data_as(C.c_void_p))
        for k in range(n):
            assert np.linalg.norm(T @ X[:, k] - w[k] * X[:, k]) <= 1e-10 * np.linalg.norm(X[:, k]) * max(1, abs(w[k]))
        # reorder: bring a scattered selection to the front, order preserved
        sel = (rng.uniform(size=n) < 0.4).astype(np.int8)
        T2, Z2 = np.asfortranarray(T.copy()), np.asfortranarray(Z.copy())
        ks = lib.dense_schur_reorder(n, T2.ctypes.data_as(C.c_void_p), Z2.ctypes.data_as(C.c_void_p),
                                     sel.tobytes())
        assert ks == sel.sum()
        assert np.allclose(np.diag(T2)[:ks], w[sel.astype(bool)], atol=1e-10)
        assert np.allclose(Z2 @ T2 @ Z2.conj().T, H, atol=1e-11 * max(1, np.abs(H).max()))
        assert np.abs(np.tril(T2, -1)).max() < 1e-300 if n > 1 else True


@pytest.mark.parametrize("which", ["LM", "SM", "LR", "SR", "LI", "SI"])
def test_iram_dense_random_matches_arpack(lib, which):
    rng = np.random.default_rng(5)
    n, nev, ncv = 120, 5, 14
    # non-normal matrix with a known spectrum whose extremal parts are well separated
    lam = (1.0 + np.arange(n)) ** 1.5 * np.exp(0.9j * np.sin(np.arange(n)))
    X = np.eye(n) + 0.1 * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    M = X @ np.diag(lam) @ np.linalg.inv(X)
    v0 = osolvers.zlarnv(n)
    ritz, vecs, st = run_iram(lib, n, lambda x: M @ x, v0, nev, ncv, 300, which=which, tol=1e-12)
    calls = [0]

    def mv(x):
        calls[0] += 1
        return M @ x

    ref, _ = eigs(LinearOperator((n, n), matvec=mv, dtype=complex), k=nev, which=which, ncv=ncv,
                  maxiter=300, tol=1e-12, v0=v0.copy())
    assert st["info"] == 0 and st["nconv"] == nev
    assert max(np.min(np.abs(ref - x)) for x in ritz) < 1e-9
    assert abs(st["n_op"] - calls[0]) <= 0.2 * calls[0] + ncv
    for k in range(nev):
        assert abs(np.linalg.norm(vecs[:, k]) - 1) < 1e-10
        assert np.linalg.norm(M @ vecs[:, k] - ritz[k] * vecs[:, k]) < 1e-8 * abs(ritz[k])


CASES = [("adiabatic_homo", 51, 15.0 + 0j, 6, 0), ("kelvin_helmholtz_cd", 51, 2.5 + 0.5j, 6, 300),
         ("magnetothermal_instabilities", 51, 0.01 + 0.04j, 15, 0),
         ("resistive_tearing", 201, 0.3 - 0.2j, 20, 0)]


@pytest.mark.parametrize("name,gridpts,sigma,nev,maxiter", CASES)
def test_iram_shift_invert_matches_arpack(lib, name, gridpts, sigma, nev, maxiter):
    """Same operator (LAPACK band LU) under our IRAM and under ARPACK: same converged set,
    similar OP*x count, eigenvalues equal to ARPACK's own accuracy."""
    so, go, xgo, fo = oeq.EQUILIBRIA[name](gridpts=gridpts)
    A, B = asm.build_matrices(so, go, xgo, fo)
    Ab, Bb = A.to_band(), B.to_band()
    n = A.n
    ncv, maxiter, tol = osolvers.arpack_defaults(n, nev, 0, maxiter, 0.0)
    lu = osolvers.BandedLU(Ab - sigma * Bb, 31, 31)
    op = lambda x: lu.solve(osolvers.banded_matvec(Bb, 31, 31, x))
    nu, vecs, st = run_iram(lib, n, op, osolvers.zlarnv(n), nev, ncv, maxiter, tol=tol)
    om_o, _, st_o = osolvers.shift_invert(Ab, Bb, 31, 31, sigma, nev, maxiter=maxiter, return_stats=True)
    assert st["nconv"] == st_o["nconv"] == nev and st["info"] == 0
    assert abs(st["n_op"] - st_o["n_op"]) <= 0.25 * st_o["n_op"] + 2 * ncv, (st, st_o["n_op"])
    omega = sigma + 1.0 / nu
    assert max(np.min(np.abs(om_o - w)) / abs(w) for w in omega) < 1e-7


def test_iram_reports_nonconvergence_like_arpack(lib):
    """maxiter reached -> info = 1, nconv < nev, converged subset still returned."""
    so, go, xgo, fo = oeq.kelvin_helmholtz_cd_eq(gridpts=51)
    A, B = asm.build_matrices(so, go, xgo, fo)
    Ab, Bb = A.to_band(), B.to_band()
    sigma, nev = 2.5 + 0.5j, 6
    lu = osolvers.BandedLU(Ab - sigma * Bb, 31, 31)
    op = lambda x: lu.solve(osolvers.banded_matvec(Bb, 31, 31, x))
    nu, vecs, st = run_iram(lib, A.n, op, osolvers.zlarnv(A.n), nev, 12, 100)
    om_o, _, st_o = osolvers.shift_invert(Ab, Bb, 31, 31, sigma, nev, return_stats=True)
    assert st["info"] == 1 and 0 < st["nconv"] < nev
    assert abs(st["nconv"] - st_o["nconv"]) <= 1
    omega = sigma + 1.0 / nu
    assert max(np.min(np.abs(om_o - w)) / abs(w) for w in omega) < 1e-8


# ---- dense_host.hpp on its own: SIMD rotations and plane-rotation generation
def _dptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("count", [0, 1, 2, 7, 40])
def test_simd_rotations_match_scalar_and_numpy(lib, count):
    lib.dense_rot_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]
    lib.dense_rot_cols.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]
    rng = np.random.default_rng(3 + count)
    ld = 43
    c, s = 0.8, 0.6 * np.exp(0.7j)
    # rows (i, i+1) of a column-major matrix over `count` columns
    M = np.asfortranarray(rng.standard_normal((ld, 45)) + 1j * rng.standard_normal((ld, 45)))
    ref = M.copy(order="F")
    x, y = ref[5, 3:3 + count].copy(), ref[6, 3:3 + count].copy()
    ref[5, 3:3 + count], ref[6, 3:3 + count] = c * x + s * y, -np.conj(s) * x + c * y
    for simd in (0, 1):
        W = M.copy(order="F")
        lib.dense_rot_rows(C.c_void_p(W.ctypes.data + 16 * (3 * ld + 5)), ld, count, c, s.real, s.imag, simd)
        assert np.abs(W - ref).max() <= 4e-16 * max(np.abs(ref).max(), 1.0), simd
    # two columns over `count` rows
    a0 = rng.standard_normal(count) + 1j * rng.standard_normal(count)
    b0 = rng.standard_normal(count) + 1j * rng.standard_normal(count)
    for simd in (0, 1):
        a, b = a0.copy(), b0.copy()
        lib.dense_rot_cols(_dptr(a), _dptr(b), count, c, s.real, s.imag, simd)
        assert np.abs(a - (c * a0 + np.conj(s) * b0)).max(initial=0.0) <= 4e-16 * 3
        assert np.abs(b - (-s * a0 + c * b0)).max(initial=0.0) <= 4e-16 * 3


@pytest.mark.parametrize("f,g", [(1.0 + 2.0j, -0.5 + 0.25j), (3e-200j, 1e-190), (0.0j, 2.0 - 1.0j), (1.5 + 0j, 0.0j),
                                 (1e160 + 1e160j, 2e159 - 1e150j)])
def test_plane_rotation_annihilates_second_entry(lib, f, g):
    """zlartg contract on both branches (unscaled fast path and the scaled one): real c, c^2 + |s|^2 = 1,
    [c s; -conj(s) c] [f; g] = [r; 0]."""
    lib.dense_lartg.argtypes = [C.c_double] * 4 + [C.c_void_p]
    out = np.zeros(5)
    lib.dense_lartg(f.real, f.imag, complex(g).real, complex(g).imag, _dptr(out))
    c, s, r = out[0], out[1] + 1j * out[2], out[3] + 1j * out[4]
    scale = max(abs(f), abs(g))
    assert abs(c * c + abs(s) ** 2 - 1.0) <= 1e-14
    assert abs(c * f + s * g - r) <= 1e-14 * scale
    assert abs(-np.conj(s) * f + c * g) <= 1e-14 * scale
