// Small dense complex kernels that run on the host once per Arnoldi restart (ncv x ncv,
// ncv ~ 40): Schur form of the upper Hessenberg projection, eigenvectors of a triangular
// matrix, Schur-form reordering, and ARPACK's sort.  They restate the published algorithms
// behind the LAPACK/ARPACK routines the reference reaches through znaupd/zneupd
// (zlahqr, ztrevc, ztrexc/ztrsen, zlartg, zsortc); LAPACK and arpack-ng are un-vendored,
// unpinned dependencies of the reference (CMakeLists.txt:98-103,
// .github/workflows/unit.yml:97-105).  Column-major storage throughout.
#pragma once

#include <algorithm>
#include <cmath>
#include <complex>
#include <limits>
#include <vector>

namespace lgpu {
namespace dense {

using cplx = std::complex<double>;

inline double cabs1(cplx z) { return std::fabs(z.real()) + std::fabs(z.imag()); }

// Plane rotation [c s; -conj(s) c] [f; g] = [r; 0] with real c (zlartg).
inline void lartg(cplx f, cplx g, double* c, cplx* s, cplx* r) {
  const double g1 = std::abs(g);
  if (g1 == 0.0) {
    *c = 1.0; *s = 0.0; *r = f;
    return;
  }
  const double f1 = std::abs(f);
  if (f1 == 0.0) {
    *c = 0.0; *s = std::conj(g) / g1; *r = g1;
    return;
  }
  const double d = std::hypot(f1, g1);
  *c = f1 / d;
  const cplx fs = f / f1;
  *s = fs * std::conj(g) / d;
  *r = fs * d;
}

// Schur decomposition of an upper Hessenberg matrix: H <- T (upper triangular), Z <- Z * U
// with U^H H U = T (single-shift QR, Wilkinson shifts, as zlahqr).  Z must be initialised by
// the caller (identity for a plain decomposition).  Returns 0, or k > 0 if the eigenvalue at
// position k - 1 failed to converge.
inline int hessenberg_schur(int n, cplx* H, int ldh, cplx* Z, int ldz, int nz, cplx* w) {
  auto h = [&](int i, int j) -> cplx& { return H[static_cast<size_t>(j) * ldh + i]; };
  auto z = [&](int i, int j) -> cplx& { return Z[static_cast<size_t>(j) * ldz + i]; };
  const double ulp = std::numeric_limits<double>::epsilon();
  const double smlnum = std::numeric_limits<double>::min() * (n / ulp);
  for (int j = 0; j + 2 < n; ++j)
    for (int i = j + 2; i < n; ++i) h(i, j) = 0.0;
  double hnorm = 0.0;
  for (int j = 0; j < n; ++j) {
    double s = 0.0;
    for (int i = 0; i <= std::min(j + 1, n - 1); ++i) s += std::abs(h(i, j));
    hnorm = std::max(hnorm, s);
  }
  const int itmax = 30 * std::max(10, n);
  int ihi = n - 1;
  int its = 0;
  while (ihi >= 0) {
    int l = ihi;
    while (l > 0) {
      const double sub = cabs1(h(l, l - 1));
      if (sub <= smlnum) break;
      double tst = cabs1(h(l - 1, l - 1)) + cabs1(h(l, l));
      if (tst == 0.0) tst = hnorm;
      if (sub <= ulp * tst) {
        // conservative deflation test (Ahues & Tisseur), as in zlahqr
        const double ab = std::max(cabs1(h(l, l - 1)), cabs1(h(l - 1, l)));
        const double ba = std::min(cabs1(h(l, l - 1)), cabs1(h(l - 1, l)));
        const double aa = std::max(cabs1(h(l, l)), cabs1(h(l - 1, l - 1) - h(l, l)));
        const double bb = std::min(cabs1(h(l, l)), cabs1(h(l - 1, l - 1) - h(l, l)));
        const double s = aa + ab;
        if (ba * (ab / s) <= std::max(smlnum, ulp * (bb * (aa / s)))) break;
      }
      --l;
    }
    if (l > 0) h(l, l - 1) = 0.0;
    if (l == ihi) {
      w[ihi] = h(ihi, ihi);
      --ihi;
      its = 0;
      continue;
    }
    if (++its > itmax) return ihi + 1;
    cplx mu;
    if (its % 10 == 0) {
      mu = h(ihi, ihi) + 0.75 * std::fabs(h(ihi, ihi - 1).real());   // exceptional shift
    } else {
      // eigenvalue of the trailing 2x2 closer to h(ihi, ihi)
      const cplx a = h(ihi - 1, ihi - 1), b = h(ihi - 1, ihi), c = h(ihi, ihi - 1), d = h(ihi, ihi);
      mu = d;
      const cplx bc = b * c;
      if (cabs1(bc) != 0.0) {
        const cplx x = 0.5 * (a - d);
        cplx y = std::sqrt(x * x + bc);
        if ((x.real() * y.real() + x.imag() * y.imag()) < 0.0) y = -y;
        mu = d - bc / (x + y);
      }
    }
    cplx x = h(l, l) - mu, y = h(l + 1, l);
    for (int k = l; k < ihi; ++k) {
      double c; cplx s, r;
      lartg(x, y, &c, &s, &r);
      if (k > l) { h(k, k - 1) = r; h(k + 1, k - 1) = 0.0; }
      for (int j = k; j < n; ++j) {
        const cplx t = c * h(k, j) + s * h(k + 1, j);
        h(k + 1, j) = -std::conj(s) * h(k, j) + c * h(k + 1, j);
        h(k, j) = t;
      }
      const int jmax = std::min(k + 2, ihi);
      for (int j = 0; j <= jmax; ++j) {
        const cplx t = c * h(j, k) + std::conj(s) * h(j, k + 1);
        h(j, k + 1) = -s * h(j, k) + c * h(j, k + 1);
        h(j, k) = t;
      }
      for (int j = 0; j < nz; ++j) {
        const cplx t = c * z(j, k) + std::conj(s) * z(j, k + 1);
        z(j, k + 1) = -s * z(j, k) + c * z(j, k + 1);
        z(j, k) = t;
      }
      if (k + 1 < ihi) { x = h(k + 1, k); y = h(k + 2, k); }
    }
  }
  return 0;
}

// Right eigenvectors of the leading m x m block of an upper triangular T: X(:, k) with
// X(k, k) = 1, X(j > k, k) = 0 (ztrevc, back substitution with the same small-pivot guard).
inline void triangular_eigvecs(int m, const cplx* T, int ldt, cplx* X, int ldx) {
  auto t = [&](int i, int j) { return T[static_cast<size_t>(j) * ldt + i]; };
  auto x = [&](int i, int j) -> cplx& { return X[static_cast<size_t>(j) * ldx + i]; };
  const double ulp = std::numeric_limits<double>::epsilon();
  const double smlnum = std::numeric_limits<double>::min() * (m / ulp);
  for (int k = m - 1; k >= 0; --k) {
    const double smin = std::max(ulp * cabs1(t(k, k)), smlnum);
    for (int j = 0; j < m; ++j) x(j, k) = 0.0;
    x(k, k) = 1.0;
    for (int j = 0; j < k; ++j) x(j, k) = -t(j, k);
    for (int j = k - 1; j >= 0; --j) {
      cplx d = t(j, j) - t(k, k);
      if (cabs1(d) < smin) d = smin;
      x(j, k) /= d;
      const cplx xj = x(j, k);
      for (int i = 0; i < j; ++i) x(i, k) -= xj * t(i, j);
    }
  }
}

// Swap the adjacent diagonal entries k, k+1 of the upper triangular T (ztrexc step).
inline void schur_swap(int n, cplx* T, int ldt, cplx* Z, int ldz, int nz, int k) {
  auto t = [&](int i, int j) -> cplx& { return T[static_cast<size_t>(j) * ldt + i]; };
  auto z = [&](int i, int j) -> cplx& { return Z[static_cast<size_t>(j) * ldz + i]; };
  const cplx t11 = t(k, k), t22 = t(k + 1, k + 1);
  double c; cplx s, r;
  lartg(t(k, k + 1), t22 - t11, &c, &s, &r);
  for (int j = k + 2; j < n; ++j) {
    const cplx a = t(k, j), b = t(k + 1, j);
    t(k, j) = c * a + s * b;
    t(k + 1, j) = c * b - std::conj(s) * a;
  }
  for (int i = 0; i < k; ++i) {
    const cplx a = t(i, k), b = t(i, k + 1);
    t(i, k) = c * a + std::conj(s) * b;
    t(i, k + 1) = c * b - s * a;
  }
  t(k, k) = t22;
  t(k + 1, k + 1) = t11;
  for (int i = 0; i < nz; ++i) {
    const cplx a = z(i, k), b = z(i, k + 1);
    z(i, k) = c * a + std::conj(s) * b;
    z(i, k + 1) = c * b - s * a;
  }
}

// Move the selected eigenvalues to the leading positions, keeping their relative order (ztrsen).
inline int schur_reorder(int n, cplx* T, int ldt, cplx* Z, int ldz, int nz,
                         const std::vector<char>& select) {
  int ks = 0;
  for (int k = 0; k < n; ++k) {
    if (!select[k]) continue;
    for (int j = k; j > ks; --j) schur_swap(n, T, ldt, Z, ldz, nz, j - 1);
    ++ks;
  }
  return ks;
}

// ARPACK zsortc: shell sort of x by `which` (LM/SM: magnitude, LR/SR: real part, LI/SI:
// imaginary part; "L*" ascending so the wanted end up last), optionally permuting y alike.
inline void sortc(const char* which, bool apply, int n, cplx* x, cplx* y) {
  const char a = which[0], b = which[1];
  auto key = [&](cplx v) {
    if (b == 'M') return std::hypot(v.real(), v.imag());
    if (b == 'R') return v.real();
    return v.imag();
  };
  const bool ascending = (a == 'L');
  for (int igap = n / 2; igap > 0; igap /= 2) {
    for (int i = igap; i < n; ++i) {
      int j = i - igap;
      while (j >= 0) {
        const double k1 = key(x[j]), k2 = key(x[j + igap]);
        const bool swap = ascending ? (k1 > k2) : (k1 < k2);
        if (!swap) break;
        std::swap(x[j], x[j + igap]);
        if (apply) std::swap(y[j], y[j + igap]);
        j -= igap;
      }
    }
  }
}

}  // namespace dense
}  // namespace lgpu
