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

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define LGPU_DENSE_AVX2 1
#endif

namespace lgpu {
namespace dense {

using cplx = std::complex<double>;

// ---- plane rotations on rows / columns of a column-major matrix -------------------------------
// rot_rows:  [x_j ; y_j] <- [c s ; -conj(s) c] [x_j ; y_j]  for the row pair (x, y) = rows (i, i+1),
//            columns j0 .. j1-1 (p points at (i, j0); the two rows are adjacent in memory)
// rot_cols:  [a_j b_j] <- [a_j b_j] [c -s ; conj(s) c]      for two columns a, b, rows 0 .. m-1
// The Schur iteration spends its time here (ncv ~ 40: everything is in L1, the loops are bound by
// instruction count), so on x86-64 with AVX2 + FMA (checked at run time) one 256-bit register
// holds the row pair of a column, resp. two consecutive rows of a column.
inline void rot_rows_scalar(cplx* p, int ld, int count, double c, cplx s) {
  for (int j = 0; j < count; ++j, p += ld) {
    const cplx x = p[0], y = p[1];
    p[0] = c * x + s * y;
    p[1] = -std::conj(s) * x + c * y;
  }
}
inline void rot_cols_scalar(cplx* a, cplx* b, int m, double c, cplx s) {
  for (int j = 0; j < m; ++j) {
    const cplx x = a[j], y = b[j];
    a[j] = c * x + std::conj(s) * y;
    b[j] = -s * x + c * y;
  }
}
#ifdef LGPU_DENSE_AVX2
__attribute__((target("avx2,fma"))) inline void rot_rows_avx2(cplx* p, int ld, int count, double c, cplx s) {
  const __m256d vc = _mm256_set1_pd(c);
  const __m256d s1 = _mm256_setr_pd(s.real(), s.real(), -s.real(), -s.real());
  const __m256d s2 = _mm256_setr_pd(-s.imag(), s.imag(), -s.imag(), s.imag());
  for (int j = 0; j < count; ++j, p += ld) {
    double* q = reinterpret_cast<double*>(p);
    const __m256d v = _mm256_loadu_pd(q);                    // [xr xi yr yi]
    const __m256d sw = _mm256_permute2f128_pd(v, v, 0x01);   // [yr yi xr xi]
    const __m256d swi = _mm256_permute_pd(sw, 0x5);          // [yi yr xi xr]
    __m256d r = _mm256_mul_pd(vc, v);
    r = _mm256_fmadd_pd(s1, sw, r);
    r = _mm256_fmadd_pd(s2, swi, r);
    _mm256_storeu_pd(q, r);
  }
}
__attribute__((target("avx2,fma"))) inline void rot_cols_avx2(cplx* a, cplx* b, int m, double c, cplx s) {
  const __m256d vc = _mm256_set1_pd(c), sr = _mm256_set1_pd(s.real());
  const __m256d si = _mm256_setr_pd(s.imag(), -s.imag(), s.imag(), -s.imag());
  double* pa = reinterpret_cast<double*>(a);
  double* pb = reinterpret_cast<double*>(b);
  int j = 0;
  for (; j + 2 <= m; j += 2) {
    const __m256d x = _mm256_loadu_pd(pa + 2 * j), y = _mm256_loadu_pd(pb + 2 * j);
    const __m256d xs = _mm256_permute_pd(x, 0x5), ys = _mm256_permute_pd(y, 0x5);   // re <-> im
    __m256d t = _mm256_mul_pd(vc, x);
    t = _mm256_fmadd_pd(sr, y, t);
    t = _mm256_fmadd_pd(si, ys, t);
    __m256d u = _mm256_mul_pd(vc, y);
    u = _mm256_fnmadd_pd(sr, x, u);
    u = _mm256_fmadd_pd(si, xs, u);
    _mm256_storeu_pd(pa + 2 * j, t);
    _mm256_storeu_pd(pb + 2 * j, u);
  }
  if (j < m) rot_cols_scalar(a + j, b + j, m - j, c, s);
}
inline bool have_avx2() {
  static const bool ok = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
  return ok;
}
#endif
inline void rot_rows(cplx* p, int ld, int count, double c, cplx s) {
#ifdef LGPU_DENSE_AVX2
  if (have_avx2()) return rot_rows_avx2(p, ld, count, c, s);
#endif
  rot_rows_scalar(p, ld, count, c, s);
}
inline void rot_cols(cplx* a, cplx* b, int m, double c, cplx s) {
#ifdef LGPU_DENSE_AVX2
  if (have_avx2()) return rot_cols_avx2(a, b, m, c, s);
#endif
  rot_cols_scalar(a, b, m, c, s);
}

inline double cabs1(cplx z) { return std::fabs(z.real()) + std::fabs(z.imag()); }

// Plane rotation [c s; -conj(s) c] [f; g] = [r; 0] with real c (zlartg).
inline void lartg(cplx f, cplx g, double* c, cplx* s, cplx* r) {
  {
    // both entries comfortably inside the range where squares neither overflow nor underflow
    // (the unscaled branch of LAPACK 3.10's zlartg): no hypot, two square roots
    const double f2 = std::norm(f), g2 = std::norm(g);
    if (f2 > 1e-140 && f2 < 1e140 && g2 > 1e-140 && g2 < 1e140) {
      const double h2 = f2 + g2;
      *c = std::sqrt(f2 / h2);
      *r = f / *c;
      *s = std::conj(g) * (f / std::sqrt(f2 * h2));
      return;
    }
  }
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
      rot_rows(&h(k, k), ldh, n - k, c, s);
      rot_cols(&h(0, k), &h(0, k + 1), std::min(k + 2, ihi) + 1, c, s);
      rot_cols(&z(0, k), &z(0, k + 1), nz, c, s);
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
