// TEST-ONLY CPU double of the device KrylovOps, used to check the host-side IRAM control
// flow (legolas_b200/csrc/iram.hpp, dense_host.hpp) against SciPy's ARPACK without a GPU.
// It is never linked into liblegolas_b200.so: the product has no CPU path.
#include <complex>
#include <cstring>
#include <vector>

#include "../../legolas_b200/csrc/iram.hpp"

using lgpu::cplx;

namespace {

typedef void (*op_fn)(const double* x_ri, double* y_ri);

class CpuOps final : public lgpu::KrylovOps {
 public:
  CpuOps(int n, int ncv, op_fn op, const cplx* resid0)
      : n_(n), ncv_(ncv), op_(op), V_(static_cast<size_t>(n) * ncv), resid_(resid0, resid0 + n),
        H_(static_cast<size_t>(ncv) * ncv), Z_() {}

  void apply(const cplx* x, cplx* y) {
    std::vector<cplx> tmp(n_);
    op_(reinterpret_cast<const double*>(x), reinterpret_cast<double*>(tmp.data()));
    std::copy(tmp.begin(), tmp.end(), y);
  }
  double norm(const std::vector<cplx>& v) {
    double s = 0;
    for (auto& z : v) s += std::norm(z);
    return std::sqrt(s);
  }
  void init_residual() override {
    apply(resid_.data(), resid_.data());
    rnorm_ = norm(resid_);
  }
  void extend(int k, int m) override {
    if (k == 0) rnorm_ = norm(resid_);
    for (int j = k; j < m; ++j) {
      cplx* vj = &V_[static_cast<size_t>(j) * n_];
      for (int i = 0; i < n_; ++i) vj[i] = resid_[i] / rnorm_;
      if (j > 0) H_[static_cast<size_t>(j - 1) * ncv_ + j] = rnorm_;
      apply(vj, resid_.data());
      for (int c = 0; c <= j; ++c) H_[static_cast<size_t>(j) * ncv_ + c] = 0.0;
      for (int pass = 0; pass < 2; ++pass) {
        std::vector<cplx> h(j + 1);
        for (int c = 0; c <= j; ++c) {
          cplx s = 0;
          const cplx* vc = &V_[static_cast<size_t>(c) * n_];
          for (int i = 0; i < n_; ++i) s += std::conj(vc[i]) * resid_[i];
          h[c] = s;
        }
        for (int c = 0; c <= j; ++c) {
          const cplx* vc = &V_[static_cast<size_t>(c) * n_];
          for (int i = 0; i < n_; ++i) resid_[i] -= vc[i] * h[c];
          H_[static_cast<size_t>(j) * ncv_ + c] += h[c];
        }
      }
      rnorm_ = norm(resid_);
    }
  }
  void fetch(int k, int m, cplx* H, int ldh, double* rnorm) override {
    for (int j = k; j < m; ++j) {
      for (int i = 0; i <= j; ++i) H[static_cast<size_t>(j) * ldh + i] = H_[static_cast<size_t>(j) * ncv_ + i];
      if (j > 0) H[static_cast<size_t>(j - 1) * ldh + j] = H_[static_cast<size_t>(j - 1) * ncv_ + j];
    }
    *rnorm = rnorm_;
  }
  void gemm(int nk, const cplx* Q, int ldq, int nc, std::vector<cplx>& out) {
    out.assign(static_cast<size_t>(n_) * nc, cplx(0.0));
    for (int c = 0; c < nc; ++c)
      for (int j = 0; j < nk; ++j) {
        const cplx q = Q[static_cast<size_t>(c) * ldq + j];
        const cplx* vj = &V_[static_cast<size_t>(j) * n_];
        cplx* oc = &out[static_cast<size_t>(c) * n_];
        for (int i = 0; i < n_; ++i) oc[i] += vj[i] * q;
      }
  }
  void compress(int kplusp, int kev, const cplx* Q, int ldq, cplx sigmak, double betak) override {
    std::vector<cplx> out;
    gemm(kplusp, Q, ldq, kev + 1, out);
    std::copy(out.begin(), out.end(), V_.begin());
    const cplx* vk = &V_[static_cast<size_t>(kev) * n_];
    for (int i = 0; i < n_; ++i) resid_[i] = sigmak * resid_[i] + betak * vk[i];
    rnorm_ = norm(resid_);
  }
  void ritz_vectors(int kplusp, int nconv, const cplx* S, int lds) override {
    gemm(kplusp, S, lds, nconv, Z_);
  }
  const std::vector<cplx>& Z() const { return Z_; }

 private:
  int n_, ncv_;
  op_fn op_;
  std::vector<cplx> V_, resid_, H_, Z_;
  double rnorm_ = 0.0;
};

}  // namespace

extern "C" {

int iram_cpu_run(int n, op_fn op, const double* resid0_ri, int nev, int ncv, int maxiter,
                 const char* which, double tol, double* ritz_ri, double* vecs_ri, int* stats) {
  lgpu::IramConfig cfg;
  cfg.nev = nev; cfg.ncv = ncv; cfg.maxiter = maxiter; cfg.tol = tol;
  cfg.which[0] = which[0]; cfg.which[1] = which[1];
  CpuOps ops(n, ncv, op, reinterpret_cast<const cplx*>(resid0_ri));
  lgpu::Iram iram;
  lgpu::IramResult res = iram.run(ops, cfg);
  for (int k = 0; k < res.nconv; ++k) {
    ritz_ri[2 * k] = res.ritz[k].real();
    ritz_ri[2 * k + 1] = res.ritz[k].imag();
  }
  if (res.nconv > 0)
    std::memcpy(vecs_ri, ops.Z().data(), sizeof(cplx) * static_cast<size_t>(n) * res.nconv);
  stats[0] = res.info; stats[1] = res.nconv; stats[2] = res.n_op; stats[3] = res.n_iter;
  return 0;
}

// H (n x n, column-major, upper Hessenberg) -> T in place, Z = Schur vectors, w = eigenvalues
int dense_hessenberg_schur(int n, double* H_ri, double* Z_ri, double* w_ri) {
  cplx* Z = reinterpret_cast<cplx*>(Z_ri);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) Z[static_cast<size_t>(j) * n + i] = i == j ? 1.0 : 0.0;
  return lgpu::dense::hessenberg_schur(n, reinterpret_cast<cplx*>(H_ri), n, Z, n, n,
                                       reinterpret_cast<cplx*>(w_ri));
}

void dense_triangular_eigvecs(int m, const double* T_ri, int ldt, double* X_ri) {
  lgpu::dense::triangular_eigvecs(m, reinterpret_cast<const cplx*>(T_ri), ldt,
                                  reinterpret_cast<cplx*>(X_ri), ldt);
}

int dense_schur_reorder(int n, double* T_ri, double* Z_ri, const char* select) {
  std::vector<char> sel(select, select + n);
  return lgpu::dense::schur_reorder(n, reinterpret_cast<cplx*>(T_ri), n,
                                    reinterpret_cast<cplx*>(Z_ri), n, n, sel);
}

// dense_host.hpp pieces on their own (column-major, interleaved complex): the AVX2 + FMA plane
// rotations against the scalar ones, and the plane-rotation generator
int dense_have_avx2() {
#ifdef LGPU_DENSE_AVX2
  return lgpu::dense::have_avx2() ? 1 : 0;
#else
  return 0;
#endif
}

void dense_rot_rows(double* p_ri, int ld, int count, double c, double s_re, double s_im, int simd) {
  cplx* p = reinterpret_cast<cplx*>(p_ri);
  if (simd) lgpu::dense::rot_rows(p, ld, count, c, cplx(s_re, s_im));
  else lgpu::dense::rot_rows_scalar(p, ld, count, c, cplx(s_re, s_im));
}

void dense_rot_cols(double* a_ri, double* b_ri, int m, double c, double s_re, double s_im, int simd) {
  cplx* a = reinterpret_cast<cplx*>(a_ri);
  cplx* b = reinterpret_cast<cplx*>(b_ri);
  if (simd) lgpu::dense::rot_cols(a, b, m, c, cplx(s_re, s_im));
  else lgpu::dense::rot_cols_scalar(a, b, m, c, cplx(s_re, s_im));
}

void dense_lartg(double f_re, double f_im, double g_re, double g_im, double* out5) {
  double c; cplx s, r;
  lgpu::dense::lartg(cplx(f_re, f_im), cplx(g_re, g_im), &c, &s, &r);
  out5[0] = c; out5[1] = s.real(); out5[2] = s.imag(); out5[3] = r.real(); out5[4] = r.imag();
}

}  // extern "C"
