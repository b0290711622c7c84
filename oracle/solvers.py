"""CPU restatement of the shift-invert Arnoldi path (and dense QR-invert for small grids).

Test infrastructure (see ``oracle/__init__.py``).  The reference delegates all
arithmetic on this path to un-vendored, *unpinned* third-party libraries:
BLAS/LAPACK (system ``liblapack-dev``, ``CMakeLists.txt:98-103``) and ARPACK
(``arpack-ng`` git HEAD, ``.github/workflows/unit.yml:97-105``).  Here the same
routines are taken from SciPy's bundled OpenBLAS (``zgbtrf/zgbtrs/zgbmv/zlarnv``)
and SciPy's ARPACK (``znaupd/zneupd`` behind ``scipy.sparse.linalg.eigs``) and are
driven exactly like the reference's call sites:

  solve_arpack_shift_invert .. src/solvers/arnoldi/smod_arpack_shift_invert.f08:15-161
  arpack_t defaults .......... src/solvers/arnoldi/mod_arpack_type.f08:74-102,194-275
  band LU / solve ............ src/solvers/mod_linear_systems.f08:67-127
  banded matvec .............. src/matrices/datastructure/mod_banded_operations.f08:18-41
  QR-invert .................. src/solvers/smod_qr_invert.f08:46-135
  solve_arpack_general ....... src/solvers/arnoldi/smod_arpack_general.f08:14-131

Parity of this file is pinned by the reference's own known answers
(``tests/unit_tests/mod_test_solvers_arpack_shift_invert.pf:16-27``) and the
``BASE_*_SI_*.dat`` baselines (see tests/test_oracle_solvers.py).
"""
from __future__ import annotations

import ctypes
import glob
import os
import time

import numpy as np
import scipy
import scipy.linalg
from scipy.linalg import blas, lapack
from scipy.sparse.linalg import ArpackNoConvergence, LinearOperator, eigs

ZLARNV_SEED = (2022, 9, 30, 179)  # mod_arpack_type.f08:206


def _openblas():
    root = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    libs = glob.glob(os.path.join(root, "libscipy_openblas*.so"))
    if not libs:
        raise OSError("scipy-bundled OpenBLAS not found")
    return ctypes.CDLL(libs[0])


def zlarnv(n: int, idist: int = 2, seed=ZLARNV_SEED) -> np.ndarray:
    """LAPACK ``zlarnv`` (mod_arpack_type.f08:194-211): the Arnoldi start vector."""
    lib = _openblas()
    iseed = (ctypes.c_int * 4)(*seed)
    out = np.empty(n, dtype=np.complex128)
    lib.scipy_zlarnv_(ctypes.byref(ctypes.c_int(idist)), iseed, ctypes.byref(ctypes.c_int(n)),
                      out.ctypes.data_as(ctypes.c_void_p))
    return out


def arpack_defaults(n, nev, ncv=0, maxiter=0, tol=0.0):
    """ncv / maxiter / tol defaults (mod_arpack_type.f08:217-275,
    src/settings/mod_solver_settings.f08:30-44)."""
    if ncv == 0:
        ncv = max(nev + 1, min(2 * nev, n))
    if maxiter == 0:
        maxiter = max(100, 10 * nev)
    if tol == 0.0:
        tol = 5.0e-15
    return ncv, maxiter, tol


class BandedLU:
    """zgbtrf factors of a band matrix given in LAPACK (kl+ku+1, n) layout."""

    def __init__(self, ab, kl, ku):
        n = ab.shape[1]
        lu_in = np.zeros((2 * kl + ku + 1, n), dtype=np.complex128, order="F")
        lu_in[kl:, :] = ab   # mod_linear_systems.f08:114-121
        self.kl, self.ku = kl, ku
        self.lu, self.ipiv, self.info = lapack.zgbtrf(lu_in, kl, ku, overwrite_ab=True)

    def solve(self, rhs):
        x, info = lapack.zgbtrs(self.lu, self.kl, self.ku, rhs, self.ipiv)
        return x


def banded_matvec(ab, kl, ku, x):
    n = ab.shape[1]
    return blas.zgbmv(n, n, kl, ku, 1.0, ab, x)


def shift_invert(A_band, B_band, kl, ku, sigma, nev, ncv=0, maxiter=0, tol=0.0, which="LM",
                 v0=None, return_stats=False):
    """Restatement of ``solve_arpack_shift_invert``.

    Returns ``(omega, vr)``; with ``return_stats`` also a dict with n_op, timings and
    the number of converged pairs (ARPACK info=1 "maxiter reached" is only a warning
    in the reference, so non-converged runs return the converged subset).
    """
    n = A_band.shape[1]
    ncv, maxiter, tol = arpack_defaults(n, nev, ncv, maxiter, tol)
    if v0 is None:
        v0 = zlarnv(n)
    stats = {"n_op": 0, "t_matvec": 0.0, "t_solve": 0.0}
    t0 = time.perf_counter()
    lu = BandedLU(A_band - sigma * B_band, kl, ku)   # :56-59
    stats["t_factor"] = time.perf_counter() - t0
    stats["lu_info"] = int(lu.info)

    def op(x):
        t1 = time.perf_counter()
        u = banded_matvec(B_band, kl, ku, x)   # :98
        t2 = time.perf_counter()
        y = lu.solve(u)                        # :99-104
        t3 = time.perf_counter()
        stats["n_op"] += 1
        stats["t_matvec"] += t2 - t1
        stats["t_solve"] += t3 - t2
        return y

    OP = LinearOperator((n, n), matvec=op, dtype=np.complex128)
    t0 = time.perf_counter()
    try:
        nu, vr = eigs(OP, k=nev, which=which, ncv=ncv, maxiter=maxiter, tol=tol, v0=v0.copy())
        nconv = nev
    except ArpackNoConvergence as exc:
        nu, vr = exc.eigenvalues, exc.eigenvectors
        nconv = len(nu)
    stats["t_iter"] = time.perf_counter() - t0
    stats["nconv"] = nconv
    omega = sigma + 1.0 / nu                   # :157
    if return_stats:
        return omega, vr, stats
    return omega, vr


def arnoldi_general(A_band, B_band, kl, ku, nev, ncv=0, maxiter=0, tol=0.0, which="LM", v0=None,
                    return_stats=False):
    """Restatement of ``solve_arpack_general`` (src/solvers/arnoldi/smod_arpack_general.f08:14-131):
    ARPACK mode 1, bmat = "I" on OP = B^-1 A; every OP*x is ``zgbmv`` with A (:88) followed by
    ``solve_linear_system_complex_banded`` = ``zgbsv`` on a fresh copy of B (:89-91,
    src/solvers/mod_linear_systems.f08:33-62) -- i.e. the reference factorises B again on every
    application; the factors are the same each time, so one ``zgbtrf`` + ``zgbtrs`` per
    application is arithmetically identical.  The Ritz values are omega themselves."""
    n = A_band.shape[1]
    ncv, maxiter, tol = arpack_defaults(n, nev, ncv, maxiter, tol)
    if v0 is None:
        v0 = zlarnv(n)
    lu = BandedLU(B_band, kl, ku)
    stats = {"n_op": 0, "lu_info": int(lu.info)}

    def op(x):
        stats["n_op"] += 1
        return lu.solve(banded_matvec(A_band, kl, ku, x))

    OP = LinearOperator((n, n), matvec=op, dtype=np.complex128)
    try:
        omega, vr = eigs(OP, k=nev, which=which, ncv=ncv, maxiter=maxiter, tol=tol, v0=v0.copy())
    except ArpackNoConvergence as exc:
        omega, vr = exc.eigenvalues, exc.eigenvectors
    stats["nconv"] = len(omega)
    if return_stats:
        return omega, vr, stats
    return omega, vr


def qr_invert(A_dense, B_dense):
    """Dense B^-1 A then zgeev (smod_qr_invert.f08:46-135); small grids only."""
    C = scipy.linalg.solve(B_dense, A_dense)
    return scipy.linalg.eigvals(C)


def _blocktri_matvec_ld(blocks_ld, x_ld):
    """NumPy fallback of oracle/extprec.c (small cases, cross-check of the C helper)."""
    d = blocks_ld.shape[-1]
    xb = x_ld.reshape(-1, d)
    y = np.einsum("bij,bj->bi", blocks_ld[:, 1], xb)
    y[1:] += np.einsum("bij,bj->bi", blocks_ld[1:, 0], xb[:-1])
    y[:-1] += np.einsum("bij,bj->bi", blocks_ld[:-1, 2], xb[1:])
    return y.reshape(-1)


class ExtendedOperator:
    """OP = (A - sigma B)^-1 B applied forward-accurately: y from a double-precision solver
    (``solve``: LAPACK zgbtrf/zgbtrs by default, or any callable b -> M^-1 b such as the device
    solve of the library under test), then iterative refinement whose residuals b - (A - sigma B) y
    and whose right-hand side b = B x are formed in 80-bit arithmetic (oracle/extprec.c) from the
    double entries of A and B, with y carried in extended precision.  The refinement stops when the
    correction no longer shrinks or is below 2^-60 relative; the accuracy of the result is
    therefore certified by the extended-precision residual, not by the solver that proposed the
    corrections.  ``stats`` records the sweeps used and the size of the last correction."""

    def __init__(self, A, B, sigma, solve=None, max_sweeps=40):
        from . import extprec

        self._gemv = extprec.gemv_ld
        self.A, self.B, self.sigma = A, B, complex(sigma)
        self.n = A.n
        if solve is None:
            kl = ku = 2 * A.d - 1
            lu = BandedLU(A.to_band() - sigma * B.to_band(), kl, ku)
            solve = lu.solve
        self.solve = solve
        self.max_sweeps = max_sweeps
        self.stats = {"n_op": 0, "sweeps_max": 0, "sweeps_total": 0, "last_correction_max": 0.0}

    def __call__(self, x):
        A, B, sigma = self.A.blocks, self.B.blocks, self.sigma
        b = self._gemv(B, None, 0.0, np.asarray(x, dtype=np.complex128).astype(np.clongdouble))
        y = self.solve(b.astype(np.complex128)).astype(np.clongdouble)
        prev = np.inf
        sweeps = 0
        rel = 0.0
        for sweeps in range(1, self.max_sweeps + 1):
            r = self._gemv(A, B, sigma, y, z_ld=b, sign=-1)
            dy = self.solve(r.astype(np.complex128))
            rel = float(np.linalg.norm(dy) / np.linalg.norm(y.astype(np.complex128)))
            if rel >= prev:            # no longer shrinking: keep y (the correction is noise)
                break
            y = y + dy.astype(np.clongdouble)
            prev = rel
            if rel <= 2.0 ** -60:
                break
        st = self.stats
        st["n_op"] += 1
        st["sweeps_max"] = max(st["sweeps_max"], sweeps)
        st["sweeps_total"] += sweeps
        st["last_correction_max"] = max(st["last_correction_max"], min(rel, prev))
        return y.astype(np.complex128)


def shift_invert_extended(A, B, sigma, nev, ncv=0, maxiter=0, tol=0.0, which="LM", v0=None,
                          solve=None, return_stats=False, max_sweeps=40):
    """Arbiter for ill-conditioned eigenvalues: the same ARPACK run as ``shift_invert``
    (reference call pattern smod_arpack_shift_invert.f08:63-157), but every OP*x is made
    forward-accurate by ``ExtendedOperator`` (A, B: oracle.assembly.BlockTriMatrix).  Used by the
    parity tests when the LAPACK-based path and the GPU path disagree beyond 1e-8: both are
    backward stable, so the one closer to this result is the better answer.
    Pinned by tests/test_oracle_golden.py: it reproduces the reference's pFUnit known answers and
    the stored shift-invert baselines to the same tolerance as ``shift_invert``."""
    n = A.n
    ncv, maxiter, tol = arpack_defaults(n, nev, ncv, maxiter, tol)
    if v0 is None:
        v0 = zlarnv(n)
    op = ExtendedOperator(A, B, sigma, solve=solve, max_sweeps=max_sweeps)
    OP = LinearOperator((n, n), matvec=op, dtype=np.complex128)
    try:
        nu, vr = eigs(OP, k=nev, which=which, ncv=ncv, maxiter=maxiter, tol=tol, v0=v0.copy())
    except ArpackNoConvergence as exc:
        nu, vr = exc.eigenvalues, exc.eigenvectors
    op.stats["nconv"] = len(nu)
    omega = sigma + 1.0 / nu
    if return_stats:
        return omega, vr, op.stats
    return omega, vr


# --------------------------------------------------------------------------- rows N1 / N4
def residuals(Ab, Bb, kl, ku, omega, vr):
    """get_residual for every eigenpair (src/dataIO/mod_output.f08:511-545):
    || A v - omega B v ||_2 / || omega v ||_2, and 0 where omega is zero by is_zero
    (|Re| <= 5e-15 and |Im| <= 5e-15, src/mod_check_values.f08:143-161)."""
    omega = np.asarray(omega, dtype=np.complex128)
    out = np.zeros(len(omega))
    for k, om in enumerate(omega):
        if abs(om.real) <= 5.0e-15 and abs(om.imag) <= 5.0e-15:
            continue
        v = np.ascontiguousarray(vr[:, k])
        y = banded_matvec(Ab, kl, ku, v) - om * banded_matvec(Bb, kl, ku, v)
        out[k] = np.linalg.norm(y) / np.linalg.norm(om * v)
    return out


def inverse_iteration(Ab, Bb, kl, ku, sigma, maxiter=0, tol=5.0e-15, start="lapack"):
    """inverse_iteration (src/solvers/smod_inverse_iteration.f08:16-205) on LAPACK band storage.

    start = "lapack": x0 solves U x0 = 1 with the U of zgbtrf, as the reference (:117-119);
    start = "solve":  x0 = (A - sigma B)^-1 1, the start vector of the device implementation (its
    factors are not LAPACK's).  B is applied as stored; the reference goes through zhbmv on the
    upper triangle, identical for the Hermitian B of every configuration without Hall terms.
    Returns (omega, x with its largest entry made real, info dict)."""
    from scipy.linalg import blas

    if maxiter == 0:
        maxiter = 100                       # :59-61
    if maxiter < 0:
        raise ValueError(f"maxiter has to be positive, but is equal to {maxiter}")
    if sigma == 0:
        raise ValueError("inverse-iteration: sigma can not be equal to zero")
    n = Ab.shape[1]
    lu = BandedLU(Ab - sigma * Bb, kl, ku)
    ones = np.ones(n, dtype=np.complex128)
    if start == "lapack":
        # ztbsv('U','N','N', n, kl+ku, LU, ld, x): rows 0..kl+ku of the factored band are U
        x = blas.ztbsv(kl + ku, np.asfortranarray(lu.lu[: kl + ku + 1]), ones, lower=0)
    else:
        x = lu.solve(ones)
        x = x / np.linalg.norm(x)
    i = 0
    converged = False
    ev = complex(sigma)
    while i <= maxiter and not converged:
        r = banded_matvec(Bb, kl, ku, x)
        s = banded_matvec(Ab, kl, ku, x)
        ev = np.vdot(x, s) / np.vdot(x, r)
        s = s - ev * r
        if np.linalg.norm(s) < abs(ev) * tol:
            converged = True
            break
        i += 1
        x = lu.solve(r)
        x = x / np.linalg.norm(x)
    im = int(np.argmax(np.abs(x)))          # idamax on abs(x): first maximum
    x = x * (np.conj(x[im]) / abs(x[im]))
    return complex(ev), x, {"iterations": i, "converged": converged}
