// Implicitly restarted Arnoldi driver (host control flow; all O(N) work goes through
// KrylovOps, i.e. the device).  Restates the algorithm of ARPACK's znaupd/znaup2/znaitr/
// zneigh/zngets/znconv/znapps and zneupd (arpack-ng; un-vendored, unpinned dependency of
// the reference) for exactly the configuration the reference uses
// (src/solvers/arnoldi/smod_arpack_shift_invert.f08:63-143,
//  src/solvers/arnoldi/mod_arpack_type.f08:74-102):
//   bmat = 'I' (standard problem OP x = nu x, Euclidean inner product), ishift = 1 (exact
//   shifts), user start vector (info = 1), rvec = .true., howmny = 'A', no sigma transform
//   in the extraction (the caller applies omega = sigma + 1/nu).
// Differences from ARPACK, by design: the Gram-Schmidt step always re-orthogonalises once
// (CGS2) instead of ARPACK's conditional DGKS pass, and a whole batch of Arnoldi steps is
// issued to the device without host synchronisation.
#pragma once

#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <limits>
#include <vector>

#include "dense_host.hpp"

namespace lgpu {

using dense::cplx;

// Device-side operations on the Krylov basis V (N x ncv), the residual vector and the
// Hessenberg columns.  Implementations: CUDA (api.cu); a CPU double exists only in tests/.
struct KrylovOps {
  virtual ~KrylovOps() = default;
  // resid <- OP * resid (ARPACK zgetv0: force the start vector into the range of OP);
  // afterwards the residual norm is available through fetch().
  virtual void init_residual() = 0;
  // Arnoldi steps j = k .. m-1:  v_j = resid/rnorm ; w = OP v_j ; CGS2 against v_0..v_j ;
  // H(0:j, j) = coefficients ; H(j, j-1) = rnorm (j > 0) ; resid = w ; rnorm = ||w||.
  virtual void extend(int k, int m) = 0;
  // Wait for the device and copy columns k..m-1 of H (incl. sub-diagonals) and rnorm.
  virtual void fetch(int k, int m, cplx* H, int ldh, double* rnorm) = 0;
  // Implicit restart: V(:, 0:kev) <- V(:, 0:kplusp) Q(:, 0:kev) (+ column kev if betak > 0),
  // resid <- sigmak * resid + betak * V(:, kev), rnorm <- ||resid||.
  virtual void compress(int kplusp, int kev, const cplx* Q, int ldq, cplx sigmak,
                        double betak) = 0;
  // Z(:, 0:nconv) = V(:, 0:kplusp) S(:, 0:nconv)
  virtual void ritz_vectors(int kplusp, int nconv, const cplx* S, int lds) = 0;
};

struct IramConfig {
  int nev = 0, ncv = 0, maxiter = 0;
  char which[2] = {'L', 'M'};
  double tol = 0.0;
};

struct IramResult {
  int info = 0;       // znaupd codes: 0 ok, 1 maxiter reached, 3 no shifts could be applied (np == 0),
                      // -8 error in the small eigen-solve, -9 zero start vector; nconv = 0 on every error
  int nconv = 0;
  int n_op = 0;
  int n_reorth = 0;
  int n_iter = 0;     // Arnoldi update iterations taken
  std::vector<cplx> ritz;     // nconv converged Ritz values nu (order of the Schur diagonal)
  std::vector<double> resid;  // their Ritz estimates ||OP x - nu x||
};

class Iram {
 public:
  IramResult run(KrylovOps& ops, const IramConfig& cfg) {
    IramResult res;
    const int nev0 = cfg.nev, kplusp = cfg.ncv, np0 = cfg.ncv - cfg.nev;
    const double eps = 0.5 * std::numeric_limits<double>::epsilon();   // dlamch('E')
    const double eps23 = std::pow(eps, 2.0 / 3.0);
    const double tol = cfg.tol > 0.0 ? cfg.tol : eps;
    const int ld = kplusp;
    H_.assign(static_cast<size_t>(ld) * ld, cplx(0.0));
    std::vector<cplx> ritz(kplusp), bounds(kplusp), ritz0(kplusp), bounds0(kplusp);
    std::vector<cplx> Q(static_cast<size_t>(ld) * ld), T(static_cast<size_t>(ld) * ld);

    using clk = std::chrono::steady_clock;
    auto since = [](clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); };
    double t_enq = 0, t_wait = 0, t_dense = 0, t_compress = 0, t_neigh = 0;
    const bool trace = std::getenv("LGPU_TRACE") != nullptr;
    ops.init_residual();
    res.n_op = 1;
    int nev = nev0, np = np0;
    int kcur = 0;
    double rnorm = 0.0;
    int nconv = 0, iter = 0;
    bool first = true;
    while (true) {
      ++iter;
      auto t0 = clk::now();
      ops.extend(kcur, kplusp);
      t_enq += since(t0);
      t0 = clk::now();
      ops.fetch(kcur, kplusp, H_.data(), ld, &rnorm);
      t_wait += since(t0);
      t0 = clk::now();
      if (first) {
        first = false;
        if (!(H0norm_ok(rnorm))) { res.info = -9; return res; }
      }
      res.n_op += kplusp - kcur;
      res.n_reorth += kplusp - kcur;
      // zneigh: Ritz values and error bounds of the current H
      {
        const auto tn = clk::now();
        if (neigh(kplusp, rnorm, T, Q, ritz, bounds, false) != 0) { res.info = -8; return res; }
        t_neigh += since(tn);
      }
      ritz0 = ritz;
      bounds0 = bounds;
      nev = nev0;
      np = np0;
      ngets(cfg.which, nev, np, ritz, bounds);
      nconv = 0;
      for (int i = 0; i < nev; ++i) {
        const double rt = std::max(eps23, std::abs(ritz[np + i]));
        if (std::abs(bounds[np + i]) <= tol * rt) ++nconv;
      }
      const int nptemp = np;
      for (int j = 0; j < nptemp; ++j)
        if (bounds[j] == cplx(0.0)) { --np; ++nev; }
      if (nconv >= nev0 || iter > cfg.maxiter || np == 0) {
        if (iter > cfg.maxiter && nconv < nev0) res.info = 1;
        if (np == 0 && nconv < nev0) res.info = 3;   // znaup2 reports 2, znaupd surfaces it as 3
        break;
      }
      if (nconv < nev0) {
        const int nevbef = nev;
        nev += std::min(nconv, np / 2);
        if (nev == 1 && kplusp >= 6) nev = kplusp / 2;
        else if (nev == 1 && kplusp > 3) nev = 2;
        np = kplusp - nev;
        if (nevbef < nev) ngets(cfg.which, nev, np, ritz, bounds);
      }
      // znapps: apply the np unwanted Ritz values as shifts, compress to a nev-step factorisation
      cplx sigmak;
      double betak;
      napps(kplusp, nev, np, ritz.data(), Q, &sigmak, &betak);
      t_dense += since(t0);
      t0 = clk::now();
      ops.compress(kplusp, nev, Q.data(), ld, sigmak, betak);
      t_compress += since(t0);
      kcur = nev;
    }
    if (trace)
      std::fprintf(stderr, "[lgpu] iram: enqueue %.2f ms, wait %.2f ms, host dense %.2f ms (Ritz values/bounds %.2f), compress %.2f ms, restarts %d\n",
                   t_enq, t_wait, t_dense, t_neigh, t_compress, iter);
    res.n_iter = iter;
    res.nconv = std::min(nconv, nev0);
    // Schur form of the final H with all of Q (ritz / bounds of the loop's last pass stay in
    // ritz0 / bounds0)
    if (res.nconv > 0 && neigh(kplusp, rnorm, T, Q, ritz, bounds, true) != 0) { res.info = -8; res.nconv = 0; return res; }
    extract(ops, cfg, kplusp, nev0, np0, res.nconv, tol, eps23, rnorm, ritz0, bounds0, T, Q, res);
    return res;
  }

 private:
  std::vector<cplx> H_;

  static bool H0norm_ok(double rnorm) { return rnorm > 0.0 && std::isfinite(rnorm); }

  cplx& h(int i, int j, int ld) { return H_[static_cast<size_t>(j) * ld + i]; }
  std::vector<cplx> qrow_, X_;

  // zneigh
  // The error bounds need only the last row of the Schur vectors, and the rotations act on the
  // rows of Q independently: inside the restart loop only that row is accumulated (1 x n matrix,
  // same arithmetic as the last row of the full accumulation); the full Q is formed once, for
  // the final H, before the Ritz vectors are extracted (full_q).
  int neigh(int n, double rnorm, std::vector<cplx>& T, std::vector<cplx>& Q,
            std::vector<cplx>& ritz, std::vector<cplx>& bounds, bool full_q) {
    const int ld = n;
    T = H_;
    qrow_.assign(n, cplx(0.0));
    qrow_[n - 1] = 1.0;
    if (full_q) {
      std::fill(Q.begin(), Q.end(), cplx(0.0));
      for (int i = 0; i < n; ++i) Q[static_cast<size_t>(i) * ld + i] = 1.0;
      if (dense::hessenberg_schur(n, T.data(), ld, Q.data(), ld, n, ritz.data()) != 0) return -8;
      for (int i = 0; i < n; ++i) qrow_[i] = Q[static_cast<size_t>(i) * ld + (n - 1)];
    } else {
      if (dense::hessenberg_schur(n, T.data(), ld, qrow_.data(), 1, 1, ritz.data()) != 0) return -8;
    }
    X_.resize(static_cast<size_t>(ld) * ld);
    dense::triangular_eigvecs(n, T.data(), ld, X_.data(), ld);
    // last component of each unit-norm eigenvector of H:  (Q X)(n-1, j) / ||Q X(:, j)||
    for (int j = 0; j < n; ++j) {
      cplx last = 0.0;
      double nrm2 = 0.0;
      for (int i = 0; i <= j; ++i) {
        last += qrow_[i] * X_[static_cast<size_t>(j) * ld + i];
        nrm2 += std::norm(X_[static_cast<size_t>(j) * ld + i]);   // Q unitary: ||Q x|| = ||x||
      }
      bounds[j] = rnorm * last / std::sqrt(nrm2);
    }
    return 0;
  }

  // zngets (ishift = 1)
  static void ngets(const char* which, int kev, int np, std::vector<cplx>& ritz,
                    std::vector<cplx>& bounds) {
    dense::sortc(which, true, kev + np, ritz.data(), bounds.data());
    // shifts with the largest Ritz estimates first
    const char sm[2] = {'S', 'M'};
    dense::sortc(sm, true, np, bounds.data(), ritz.data());
  }

  // znapps on the host copy of H; returns Q (kplusp x kplusp), sigmak = Q(kplusp-1, kev-1),
  // betak = H(kev, kev-1).
  void napps(int kplusp, int kev, int np, const cplx* shift, std::vector<cplx>& Q, cplx* sigmak,
             double* betak) {
    const int ld = kplusp;
    auto q = [&](int i, int j) -> cplx& { return Q[static_cast<size_t>(j) * ld + i]; };
    const double ulp = std::numeric_limits<double>::epsilon();
    const double smlnum = std::numeric_limits<double>::min() * (kplusp / ulp);
    std::fill(Q.begin(), Q.end(), cplx(0.0));
    for (int i = 0; i < kplusp; ++i) q(i, i) = 1.0;
    auto hnorm1 = [&](int n) {
      double best = 0.0;
      for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int i = 0; i <= std::min(j + 1, n - 1); ++i) s += std::abs(h(i, j, ld));
        best = std::max(best, s);
      }
      return best;
    };
    for (int jj = 0; jj < np; ++jj) {
      const cplx sigma = shift[jj];
      int istart = 0;
      while (istart < kplusp) {
        int iend = kplusp - 1;
        for (int i = istart; i < kplusp - 1; ++i) {
          double tst1 = dense::cabs1(h(i, i, ld)) + dense::cabs1(h(i + 1, i + 1, ld));
          if (tst1 == 0.0) tst1 = hnorm1(kplusp - jj);
          if (std::fabs(h(i + 1, i, ld).real()) <= std::max(ulp * tst1, smlnum)) {
            iend = i;
            h(i + 1, i, ld) = 0.0;
            break;
          }
        }
        if (istart < iend) {
          cplx f = h(istart, istart, ld) - sigma, g = h(istart + 1, istart, ld);
          for (int i = istart; i < iend; ++i) {
            double c; cplx s, r;
            dense::lartg(f, g, &c, &s, &r);
            if (i > istart) { h(i, i - 1, ld) = r; h(i + 1, i - 1, ld) = 0.0; }
            dense::rot_rows(&h(i, i, ld), ld, kplusp - i, c, s);
            dense::rot_cols(&h(0, i, ld), &h(0, i + 1, ld), std::min(i + 2, iend) + 1, c, s);
            dense::rot_cols(&q(0, i), &q(0, i + 1), std::min(i + jj + 1, kplusp - 1) + 1, c, s);
            if (i < iend - 1) { f = h(i + 1, i, ld); g = h(i + 2, i, ld); }
          }
        }
        istart = iend + 1;
      }
    }
    // make the leading sub-diagonals real and non-negative
    for (int j = 0; j < kev; ++j) {
      const cplx sub = h(j + 1, j, ld);
      if (sub.real() < 0.0 || sub.imag() != 0.0) {
        const cplx t = sub / std::abs(sub);
        for (int c = j; c < kplusp; ++c) h(j + 1, c, ld) *= std::conj(t);
        for (int r = 0; r <= std::min(j + 2, kplusp - 1); ++r) h(r, j + 1, ld) *= t;
        for (int r = 0; r <= std::min(j + np + 1, kplusp - 1); ++r) q(r, j + 1) *= t;
        h(j + 1, j, ld) = cplx(h(j + 1, j, ld).real(), 0.0);
      }
    }
    for (int i = 0; i < kev; ++i) {
      double tst1 = dense::cabs1(h(i, i, ld)) + dense::cabs1(h(i + 1, i + 1, ld));
      if (tst1 == 0.0) tst1 = hnorm1(kev);
      if (h(i + 1, i, ld).real() <= std::max(ulp * tst1, smlnum)) h(i + 1, i, ld) = 0.0;
    }
    *sigmak = q(kplusp - 1, kev - 1);
    *betak = h(kev, kev - 1, ld).real();
    // the compressed factorisation keeps the leading kev x kev block; later columns are
    // rebuilt by the next extend()
    for (int j = kev; j < kplusp; ++j)
      for (int i = 0; i < kplusp; ++i) h(i, j, ld) = 0.0;
    for (int j = 0; j < kev; ++j)
      for (int i = kev + 1; i < kplusp; ++i) h(i, j, ld) = 0.0;
  }

  // zneupd (rvec = .true., howmny = 'A', type REGULR): converged wanted Ritz pairs.
  void extract(KrylovOps& ops, const IramConfig& cfg, int ncv, int nev, int np, int nconv,
               double tol, double eps23, double rnorm, const std::vector<cplx>& ritz0,
               const std::vector<cplx>& bounds0, std::vector<cplx>& T, std::vector<cplx>& Q,
               IramResult& res) {
    if (nconv <= 0) return;
    const int ld = ncv;
    // which Ritz values (in the order of the Schur diagonal) are wanted and converged
    std::vector<cplx> rz = ritz0, idx(ncv);
    for (int j = 0; j < ncv; ++j) idx[j] = cplx(static_cast<double>(j), 0.0);
    dense::sortc(cfg.which, true, ncv, rz.data(), idx.data());   // zngets with ishift = 0
    std::vector<char> select(ncv, 0);
    int numcnv = 0;
    for (int j = 0; j < ncv; ++j) {
      const int pos = ncv - 1 - j;
      const int jj = static_cast<int>(idx[pos].real());
      const double rt = std::max(eps23, std::abs(rz[pos]));
      if (numcnv < nconv && std::abs(bounds0[jj]) <= tol * rt) {
        select[jj] = 1;
        ++numcnv;
      }
    }
    (void)nev; (void)np;
    nconv = numcnv;
    // T, Q still hold the Schur form of the final H (from neigh)
    dense::schur_reorder(ncv, T.data(), ld, Q.data(), ld, ncv, select);
    std::vector<cplx> X(static_cast<size_t>(ld) * ld, cplx(0.0));
    dense::triangular_eigvecs(nconv, T.data(), ld, X.data(), ld);
    std::vector<cplx> S(static_cast<size_t>(ld) * nconv, cplx(0.0));
    res.ritz.resize(nconv);
    res.resid.resize(nconv);
    for (int k = 0; k < nconv; ++k) {
      double nrm2 = 0.0;
      for (int i = 0; i <= k; ++i) nrm2 += std::norm(X[static_cast<size_t>(k) * ld + i]);
      const double inv = 1.0 / std::sqrt(nrm2);
      cplx last = 0.0;
      for (int r = 0; r < ncv; ++r) {
        cplx s = 0.0;
        for (int i = 0; i <= k; ++i)
          s += Q[static_cast<size_t>(i) * ld + r] * X[static_cast<size_t>(k) * ld + i];
        S[static_cast<size_t>(k) * ld + r] = s * inv;
        if (r == ncv - 1) last = s * inv;
      }
      res.ritz[k] = T[static_cast<size_t>(k) * ld + k];
      res.resid[k] = rnorm * std::abs(last);
    }
    res.nconv = nconv;
    ops.ritz_vectors(ncv, nconv, S.data(), ld);
  }
};

}  // namespace lgpu
