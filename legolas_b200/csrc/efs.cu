// N1 — see efs.cuh.
#include "efs.cuh"

namespace lgpu {
namespace {

// src/mod_spline_functions.f08:23-37 and :60-76, same expression order
__device__ __forceinline__ void quadratic_factors(double r, double lo, double hi, double h[4]) {
  h[0] = 4.0 * (r - lo) * (hi - r) / ((hi - lo) * (hi - lo));
  h[1] = 0.0;
  h[2] = (2.0 * r - hi - lo) * (r - lo) / ((hi - lo) * (hi - lo));
  h[3] = (2.0 * r - hi - lo) * (r - hi) / ((hi - lo) * (hi - lo));
}
__device__ __forceinline__ void cubic_factors(double r, double lo, double hi, double h[4]) {
  const double a = (r - lo) / (hi - lo), b = (hi - r) / (hi - lo);
  h[0] = 3.0 * (a * a) - 2.0 * (a * a * a);
  h[1] = 3.0 * (b * b) - 2.0 * (b * b * b);
  h[2] = (r - hi) * (a * a);
  h[3] = (r - lo) * (b * b);
}

// one thread per (eigenfunction-grid point, selected eigenvector); the 8 variables in a loop
__global__ void __launch_bounds__(256)
ef_kernel(int gridpts, int geometry, const double* __restrict__ grid, const cd* __restrict__ vr, size_t ld, int nsel,
          const int32_t* __restrict__ idxs, cd* __restrict__ out) {
  const int npts = 2 * gridpts - 1;
  const int e = blockIdx.x * 256 + threadIdx.x;
  const int s = blockIdx.y;
  if (e >= npts) return;
  const int g = e == 0 ? 0 : (e - 1) >> 1;             // interval that owns the point
  const double lo = grid[g], hi = grid[g + 1];
  const double x = (e & 1) ? 0.5 * (lo + hi) : (e == 0 ? lo : hi);
  const double eps = geometry == 1 ? x : 1.0;
  double hq[4], hc[4];
  quadratic_factors(x, lo, hi, hq);
  cubic_factors(x, lo, hi, hc);
  const cd* v = vr + static_cast<size_t>(idxs[s]) * ld + static_cast<size_t>(g) * BLK;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const bool cubic = p == 1 || p == 6 || p == 7;     // v1, a2, a3
    const double* h = cubic ? hc : hq;
    // eigenvector(idx) h(2) + eigenvector(idx+1) h(4) + eigenvector(idx+dim) h(1) + eigenvector(idx+dim+1) h(3)
    cd f = v[2 * p] * h[1] + v[2 * p + 1] * h[3] + v[2 * p + BLK] * h[0] + v[2 * p + BLK + 1] * h[2];
    // retransform: true divisions, as the reference (ef / eps, ef / (eps * i), ef / i)
    if (p == 0 || p == 3 || p == 4 || p == 6) f = cd{f.x / eps, f.y / eps};           // rho, v3, T, a2
    else if (p == 1) f = cd{f.y / eps, -f.x / eps};                                   // v1: / (i eps)
    else if (p == 5) f = cd{f.y, -f.x};                                               // a1: / i
    out[(static_cast<size_t>(p) * nsel + s) * npts + e] = f;
  }
}

}  // namespace

void assemble_eigenfunctions(int gridpts, int geometry, const double* grid, const cd* vr, size_t ld, int nsel,
                             const int32_t* idxs_dev, cd* out, cudaStream_t stream, LaunchLog* log) {
  if (nsel <= 0) return;
  const int npts = 2 * gridpts - 1;
  log->begin(LK_OTHER, 16.0 * (static_cast<double>(gridpts) * BLK + 8.0 * npts) * nsel);
  ef_kernel<<<dim3((npts + 255) / 256, nsel), 256, 0, stream>>>(gridpts, geometry, grid, vr, ld, nsel, idxs_dev, out);
  log->end();
  log->launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace lgpu
