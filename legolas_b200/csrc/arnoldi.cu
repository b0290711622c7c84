// K3b / K3c — see arnoldi.cuh.
//
// All kernels here are pure streaming kernels over the N x ncv basis (HBM-bound): every
// global access is a 16-byte complex load/store with consecutive lanes on consecutive rows
// (the basis is column-major, like ARPACK's V), reductions are two-level (CTA partials, then
// the last CTA to finish sums the partials in a fixed order), so results are deterministic.
#include "arnoldi.cuh"

namespace lgpu {
namespace {

constexpr int PSTRIDE = KRYLOV_MAXCOL + 1;

__device__ __forceinline__ cd ldg_cd(const cd* p) {
  const double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return cd{v.x, v.y};
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// true for exactly one CTA: the last one to arrive; its view of global memory then contains
// every other CTA's partials.
__device__ __forceinline__ bool last_block_done(unsigned int* ticket, bool wrote_partials) {
  __shared__ bool is_last;
  if (wrote_partials) __threadfence();   // only the threads that published partials need the fence
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// h = V(:, 0:ncols)^H w.  One thread owns two rows of the tile; columns are processed in
// chunks of 8 so that 16 independent 16-byte loads per thread are in flight, then the 8 partial
// dot products are reduced across the CTA (shuffle tree + shared memory).
constexpr int DOT_CHUNK = 8;

__global__ void __launch_bounds__(256)
krylov_dots_kernel(int n, const cd* __restrict__ V, int ldv, int ncols, const cd* __restrict__ w,
                   cd* __restrict__ partial, cd* __restrict__ hwork, cd* Hcol, int accumulate,
                   unsigned int* ticket) {
  __shared__ cd red[8][DOT_CHUNK];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * KRYLOV_TILE;
  const int i0 = row0 + tid, i1 = row0 + 256 + tid;
  const bool ok0 = i0 < n, ok1 = i1 < n;
  const cd w0 = ok0 ? w[i0] : cd{0.0, 0.0}, w1 = ok1 ? w[i1] : cd{0.0, 0.0};
  for (int cb = 0; cb < ncols; cb += DOT_CHUNK) {
    cd a0[DOT_CHUNK], a1[DOT_CHUNK];
#pragma unroll
    for (int c = 0; c < DOT_CHUNK; ++c) {
      const bool okc = cb + c < ncols;
      const cd* col = V + static_cast<size_t>(cb + c) * ldv;
      a0[c] = (okc && ok0) ? ldg_cd(col + i0) : cd{0.0, 0.0};
      a1[c] = (okc && ok1) ? ldg_cd(col + i1) : cd{0.0, 0.0};
    }
#pragma unroll
    for (int c = 0; c < DOT_CHUNK; ++c) {
      cd acc{0.0, 0.0};
      cfmac(acc, a0[c], w0);
      cfmac(acc, a1[c], w1);
      acc.x = warp_sum(acc.x);
      acc.y = warp_sum(acc.y);
      if (lane == 0) red[warp][c] = acc;
    }
    __syncthreads();
    if (tid < DOT_CHUNK && cb + tid < ncols) {
      cd s{0.0, 0.0};
#pragma unroll
      for (int k = 0; k < 8; ++k) s += red[k][tid];
      partial[static_cast<size_t>(blockIdx.x) * PSTRIDE + cb + tid] = s;
    }
    __syncthreads();
  }
  if (last_block_done(ticket, tid < DOT_CHUNK)) {
    // one warp per column: lanes stride over the CTA partials, fixed-order shuffle tree
    for (int c = warp; c < ncols; c += 8) {
      cd s{0.0, 0.0};
      for (unsigned int b = lane; b < gridDim.x; b += 32) s += partial[static_cast<size_t>(b) * PSTRIDE + c];
      s.x = warp_sum(s.x);
      s.y = warp_sum(s.y);
      if (lane == 0) {
        hwork[c] = s;
        if (Hcol) Hcol[c] = accumulate ? Hcol[c] + s : s;
      }
    }
    if (tid == 0) *ticket = 0u;
  }
}

// w -= V hwork (STORE) and ||w||^2 ; with ncols == 0 it is a plain norm.
__global__ void __launch_bounds__(256)
krylov_update_kernel(int n, const cd* __restrict__ V, int ldv, int ncols, cd* __restrict__ w,
                     const cd* __restrict__ hwork, cd* __restrict__ partial, double* scal,
                     unsigned int* ticket) {
  __shared__ cd hs[KRYLOV_MAXCOL];
  __shared__ double red[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = tid; c < ncols; c += 256) hs[c] = hwork[c];
  __syncthreads();
  const int row0 = blockIdx.x * KRYLOV_TILE;
  double nrm = 0.0;
#pragma unroll
  for (int rr = 0; rr < KRYLOV_TILE / 256; ++rr) {
    const int i = row0 + rr * 256 + tid;
    if (i < n) {
      cd acc = w[i];
#pragma unroll 4
      for (int c = 0; c < ncols; ++c) cfms(acc, ldg_cd(V + static_cast<size_t>(c) * ldv + i), hs[c]);
      if (ncols > 0) w[i] = acc;
      nrm += abs2(acc);
    }
  }
  nrm = warp_sum(nrm);
  if (lane == 0) red[warp] = nrm;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int k = 0; k < 8; ++k) s += red[k];
    partial[static_cast<size_t>(blockIdx.x) * PSTRIDE + KRYLOV_MAXCOL] = cd{s, 0.0};
  }
  if (last_block_done(ticket, tid == 0)) {
    if (warp == 0) {
      double s = 0.0;
      for (unsigned int b = lane; b < gridDim.x; b += 32)
        s += partial[static_cast<size_t>(b) * PSTRIDE + KRYLOV_MAXCOL].x;
      s = warp_sum(s);
      if (lane == 0) {
        scal[0] = sqrt(s);
        *ticket = 0u;
      }
    }
  }
}

__global__ void __launch_bounds__(256)
krylov_scale_kernel(int n, const cd* __restrict__ w, cd* __restrict__ vout,
                    const double* __restrict__ scal, cd* hsub) {
  const double rnorm = scal[0];
  const double inv = 1.0 / rnorm;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) vout[i] = w[i] * inv;
  if (hsub && i == 0) *hsub = cd{rnorm, 0.0};
}

__global__ void __launch_bounds__(256)
basis_gemm_kernel(int n, const cd* __restrict__ V, int ldv, int nk, const cd* __restrict__ Q,
                  int ldq, int nc, cd* Out, int ldo) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* Qs = reinterpret_cast<cd*>(smem_raw);   // [nc][nk]
  cd* tile = Qs + nk * nc;                    // [nk][32]
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * 32;
  for (int e = tid; e < nk * nc; e += 256) {
    const int j = e % nk, c = e / nk;
    Qs[e] = Q[static_cast<size_t>(c) * ldq + j];
  }
  for (int e = tid; e < nk * 32; e += 256) {
    const int j = e >> 5, r = e & 31;
    tile[e] = (row0 + r < n) ? V[static_cast<size_t>(j) * ldv + row0 + r] : cd{0.0, 0.0};
  }
  __syncthreads();
  const int r = tid & 31, cg = tid >> 5;
  if (row0 + r >= n) return;
  for (int c = cg; c < nc; c += 8) {
    cd acc{0.0, 0.0};
    const cd* q = Qs + c * nk;
    for (int j = 0; j < nk; ++j) cfma(acc, tile[j * 32 + r], q[j]);
    Out[static_cast<size_t>(c) * ldo + row0 + r] = acc;
  }
}

__global__ void __launch_bounds__(256)
vec_axpby_kernel(int n, cd a, cd* __restrict__ r, cd b, const cd* __restrict__ v) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) r[i] = a * r[i] + b * v[i];
}

int tiles(int n) { return (n + KRYLOV_TILE - 1) / KRYLOV_TILE; }

}  // namespace

void krylov_dots(int n, const cd* V, int ldv, int ncols, const cd* w, const KrylovWork& work,
                 cd* Hcol, int accumulate, cudaStream_t stream, LaunchLog* log) {
  log->begin(LK_DOTS, 16.0 * n * (ncols + 1));
  krylov_dots_kernel<<<tiles(n), 256, 0, stream>>>(n, V, ldv, ncols, w, work.partial, work.hwork,
                                                   Hcol, accumulate, work.ticket);
  log->end();
  log->launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

void krylov_update(int n, const cd* V, int ldv, int ncols, cd* w, const KrylovWork& work,
                   cudaStream_t stream, LaunchLog* log) {
  log->begin(LK_UPDATE, 16.0 * n * (ncols + 2));
  krylov_update_kernel<<<tiles(n), 256, 0, stream>>>(n, V, ldv, ncols, w, work.hwork,
                                                     work.partial, work.scal, work.ticket);
  log->end();
  log->launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

void krylov_norm(int n, const cd* w, const KrylovWork& work, cudaStream_t stream,
                 LaunchLog* log) {
  log->begin(LK_UPDATE, 16.0 * n);
  krylov_update_kernel<<<tiles(n), 256, 0, stream>>>(n, nullptr, 0, 0, const_cast<cd*>(w),
                                                     work.hwork, work.partial, work.scal,
                                                     work.ticket);
  log->end();
  log->launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

void krylov_scale(int n, const cd* w, cd* vout, const KrylovWork& work, cd* hsub,
                  cudaStream_t stream, LaunchLog* log) {
  log->begin(LK_SCALE, 32.0 * n);
  krylov_scale_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, w, vout, work.scal, hsub);
  log->end();
  log->launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

void basis_gemm(int n, const cd* V, int ldv, int nk, const cd* Q, int ldq, int nc, cd* Out,
                int ldo, cudaStream_t stream, LaunchLog* log) {
  const size_t smem = sizeof(cd) * (static_cast<size_t>(nk) * nc + static_cast<size_t>(nk) * 32);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    CUDA_CHECK(cudaFuncSetAttribute(basis_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem)));
    configured = smem;
  }
  log->begin(LK_GEMM, 16.0 * n * (nk + nc));
  basis_gemm_kernel<<<(n + 31) / 32, 256, smem, stream>>>(n, V, ldv, nk, Q, ldq, nc, Out, ldo);
  log->end();
  log->launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

void vec_axpby(int n, cd a, cd* r, cd b, const cd* v, cudaStream_t stream, LaunchLog* log) {
  log->begin(LK_OTHER, 48.0 * n);
  vec_axpby_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, a, r, b, v);
  log->end();
  log->launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace lgpu
