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

constexpr int DOT_CHUNK = 8;

// Thread mapping of the dot / update kernels: one CTA per tile with T / 2 threads, every
// thread owns exactly the two rows tid and tid + T/2 of the tile (equal work per thread).

// h = V(:, 0:ncols)^H w.  Columns are processed in chunks of 8: 16 independent 16-byte loads
// per thread in flight, 8 partial dot products per thread, reduced through shared memory
// (warp c sums column c).
__global__ void __launch_bounds__(640)
krylov_dots_kernel(BasisLayout L, const cd* __restrict__ V, int ncols, const cd* __restrict__ w,
                   cd* __restrict__ partial, cd* __restrict__ hwork, cd* Hcol, int accumulate,
                   unsigned int* ticket) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* red = reinterpret_cast<cd*>(smem_raw);   // [DOT_CHUNK][blockDim.x]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthr = blockDim.x, nwarps = nthr >> 5;
  const int row0 = blockIdx.x * L.T;
  const int rows = min(L.T, L.n - row0);
  const cd* tile = V + static_cast<size_t>(blockIdx.x) * L.ncv * L.T;
  const int i0 = tid, i1 = tid + nthr;
  const bool ok0 = i0 < rows, ok1 = i1 < rows;
  const cd w0 = ok0 ? w[row0 + i0] : cd{0.0, 0.0}, w1 = ok1 ? w[row0 + i1] : cd{0.0, 0.0};
  for (int cb = 0; cb < ncols; cb += DOT_CHUNK) {
    cd a0[DOT_CHUNK], a1[DOT_CHUNK];
#pragma unroll
    for (int c = 0; c < DOT_CHUNK; ++c) {
      const bool okc = cb + c < ncols;
      const cd* col = tile + static_cast<size_t>(cb + c) * L.T;
      a0[c] = (okc && ok0) ? ldg_cd(col + i0) : cd{0.0, 0.0};
      a1[c] = (okc && ok1) ? ldg_cd(col + i1) : cd{0.0, 0.0};
    }
#pragma unroll
    for (int c = 0; c < DOT_CHUNK; ++c) {
      cd acc{0.0, 0.0};
      cfmac(acc, a0[c], w0);
      cfmac(acc, a1[c], w1);
      red[c * nthr + tid] = acc;
    }
    __syncthreads();
    if (warp < DOT_CHUNK && cb + warp < ncols) {
      cd s{0.0, 0.0};
      for (int k = lane; k < nthr; k += 32) s += red[warp * nthr + k];
      s.x = warp_sum(s.x);
      s.y = warp_sum(s.y);
      if (lane == 0) partial[static_cast<size_t>(blockIdx.x) * PSTRIDE + cb + warp] = s;
    }
    if (nwarps < DOT_CHUNK) {   // small tiles: fewer warps than columns in a chunk
      for (int c = nwarps + warp; c < DOT_CHUNK && cb + c < ncols; c += nwarps) {
        cd s{0.0, 0.0};
        for (int k = lane; k < nthr; k += 32) s += red[c * nthr + k];
        s.x = warp_sum(s.x);
        s.y = warp_sum(s.y);
        if (lane == 0) partial[static_cast<size_t>(blockIdx.x) * PSTRIDE + cb + c] = s;
      }
    }
    __syncthreads();
  }
  if (last_block_done(ticket, lane == 0)) {
    // one warp per column: lanes stride over the CTA partials, fixed-order shuffle tree
    for (int c = warp; c < ncols; c += nwarps) {
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
__global__ void __launch_bounds__(640)
krylov_update_kernel(BasisLayout L, const cd* __restrict__ V, int ncols, cd* __restrict__ w,
                     const cd* __restrict__ hwork, cd* __restrict__ partial, double* scal,
                     unsigned int* ticket) {
  __shared__ cd hs[KRYLOV_MAXCOL];
  __shared__ double red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthr = blockDim.x, nwarps = nthr >> 5;
  for (int c = tid; c < ncols; c += nthr) hs[c] = hwork[c];
  __syncthreads();
  const int row0 = blockIdx.x * L.T;
  const int rows = min(L.T, L.n - row0);
  const cd* tile = V + static_cast<size_t>(blockIdx.x) * L.ncv * L.T;
  const int i0 = tid, i1 = tid + nthr;
  const bool ok0 = i0 < rows, ok1 = i1 < rows;
  cd acc0 = ok0 ? w[row0 + i0] : cd{0.0, 0.0}, acc1 = ok1 ? w[row0 + i1] : cd{0.0, 0.0};
  for (int cb = 0; cb < ncols; cb += DOT_CHUNK) {
    cd a0[DOT_CHUNK], a1[DOT_CHUNK];
#pragma unroll
    for (int c = 0; c < DOT_CHUNK; ++c) {
      const bool okc = cb + c < ncols;
      const cd* col = tile + static_cast<size_t>(cb + c) * L.T;
      a0[c] = (okc && ok0) ? ldg_cd(col + i0) : cd{0.0, 0.0};
      a1[c] = (okc && ok1) ? ldg_cd(col + i1) : cd{0.0, 0.0};
    }
#pragma unroll
    for (int c = 0; c < DOT_CHUNK; ++c) {
      const cd h = cb + c < ncols ? hs[cb + c] : cd{0.0, 0.0};
      cfms(acc0, a0[c], h);
      cfms(acc1, a1[c], h);
    }
  }
  double nrm = 0.0;
  if (ok0) { if (ncols > 0) w[row0 + i0] = acc0; nrm += abs2(acc0); }
  if (ok1) { if (ncols > 0) w[row0 + i1] = acc1; nrm += abs2(acc1); }
  nrm = warp_sum(nrm);
  if (lane == 0) red[warp] = nrm;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int k = 0; k < nwarps; ++k) s += red[k];
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

__device__ __forceinline__ size_t basis_off(const BasisLayout& L, int i, int c) {
  const int t = i / L.T;
  return (static_cast<size_t>(t) * L.ncv + c) * L.T + (i - t * L.T);
}

__global__ void __launch_bounds__(256)
krylov_scale_kernel(BasisLayout L, const cd* __restrict__ w, cd* __restrict__ V, int col,
                    cd* __restrict__ vplain, const double* __restrict__ scal, cd* hsub) {
  const double rnorm = scal[0];
  const double inv = 1.0 / rnorm;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < L.n) {
    const cd v = w[i] * inv;
    V[basis_off(L, i, col)] = v;
    vplain[i] = v;
  }
  if (hsub && i == 0) *hsub = cd{rnorm, 0.0};
}

// 32 consecutive rows (inside one tile: T is a multiple of 32) per CTA.
__global__ void __launch_bounds__(256)
basis_gemm_kernel(BasisLayout L, const cd* __restrict__ V, int nk, const cd* __restrict__ Q,
                  int ldq, int nc, cd* Out, int out_plain_ld) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* Qs = reinterpret_cast<cd*>(smem_raw);   // [nc][nk]
  cd* tile = Qs + nk * nc;                    // [nk][32]
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * 32;
  const int n = L.n;
  const size_t base = basis_off(L, row0, 0);  // column c of these rows: base + c * T
  for (int e = tid; e < nk * nc; e += 256) {
    const int j = e % nk, c = e / nk;
    Qs[e] = Q[static_cast<size_t>(c) * ldq + j];
  }
  for (int e = tid; e < nk * 32; e += 256) {
    const int j = e >> 5, r = e & 31;
    tile[e] = (row0 + r < n) ? V[base + static_cast<size_t>(j) * L.T + r] : cd{0.0, 0.0};
  }
  __syncthreads();
  const int r = tid & 31, cg = tid >> 5;
  if (row0 + r >= n) return;
  for (int c = cg; c < nc; c += 8) {
    cd acc{0.0, 0.0};
    const cd* q = Qs + c * nk;
    for (int j = 0; j < nk; ++j) cfma(acc, tile[j * 32 + r], q[j]);
    if (out_plain_ld > 0) Out[static_cast<size_t>(c) * out_plain_ld + row0 + r] = acc;
    else Out[base + static_cast<size_t>(c) * L.T + r] = acc;
  }
}

__global__ void __launch_bounds__(256)
vec_axpby_basis_kernel(BasisLayout L, cd a, cd* __restrict__ r, cd b, const cd* __restrict__ V,
                       int col) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < L.n) r[i] = a * r[i] + b * V[basis_off(L, i, col)];
}

__global__ void __launch_bounds__(256)
vec_axpby_kernel(int n, cd a, cd* __restrict__ r, cd b, const cd* __restrict__ v) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) r[i] = a * r[i] + b * v[i];
}

}  // namespace

BasisLayout make_basis_layout(int n, int ncv) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  BasisLayout L{};
  L.n = n;
  L.ncv = ncv;
  int T = (n + sms - 1) / sms;             // one tile (= one CTA of T/2 threads) per SM
  T = ((T + 63) / 64) * 64;
  if (T > KRYLOV_MAX_T) T = KRYLOV_MAX_T;
  if (T < 64) T = 64;
  L.T = T;
  L.ntiles = (n + T - 1) / T;
  return L;
}

void krylov_dots(const BasisLayout& L, const cd* V, int ncols, const cd* w, const KrylovWork& work,
                 cd* Hcol, int accumulate, cudaStream_t stream, LaunchLog* log) {
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(krylov_dots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sizeof(cd) * DOT_CHUNK * 640)));
    configured = true;
  }
  log->begin(LK_DOTS, 16.0 * L.n * (ncols + 1));
  const int nthr = L.T / 2;
  krylov_dots_kernel<<<L.ntiles, nthr, sizeof(cd) * DOT_CHUNK * nthr, stream>>>(L, V, ncols, w, work.partial, work.hwork, Hcol,
                                                  accumulate, work.ticket);
  log->end();
  log->launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

void krylov_update(const BasisLayout& L, const cd* V, int ncols, cd* w, const KrylovWork& work,
                   cudaStream_t stream, LaunchLog* log) {
  log->begin(LK_UPDATE, 16.0 * L.n * (ncols + 2));
  krylov_update_kernel<<<L.ntiles, L.T / 2, 0, stream>>>(L, V, ncols, w, work.hwork, work.partial,
                                                    work.scal, work.ticket);
  log->end();
  log->launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

void krylov_scale(const BasisLayout& L, const cd* w, cd* V, int col, cd* vplain,
                  const KrylovWork& work, cd* hsub, cudaStream_t stream, LaunchLog* log) {
  log->begin(LK_SCALE, 48.0 * L.n);
  krylov_scale_kernel<<<(L.n + 255) / 256, 256, 0, stream>>>(L, w, V, col, vplain, work.scal, hsub);
  log->end();
  log->launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

void basis_gemm(const BasisLayout& L, const cd* V, int nk, const cd* Q, int ldq, int nc, cd* Out,
                int out_plain_ld, cudaStream_t stream, LaunchLog* log) {
  const size_t smem = sizeof(cd) * (static_cast<size_t>(nk) * nc + static_cast<size_t>(nk) * 32);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    CUDA_CHECK(cudaFuncSetAttribute(basis_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem)));
    configured = smem;
  }
  log->begin(LK_GEMM, 16.0 * L.n * (nk + nc));
  basis_gemm_kernel<<<(L.n + 31) / 32, 256, smem, stream>>>(L, V, nk, Q, ldq, nc, Out, out_plain_ld);
  log->end();
  log->launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

void vec_axpby_basis(const BasisLayout& L, cd a, cd* r, cd b, const cd* V, int col,
                     cudaStream_t stream, LaunchLog* log) {
  log->begin(LK_OTHER, 48.0 * L.n);
  vec_axpby_basis_kernel<<<(L.n + 255) / 256, 256, 0, stream>>>(L, a, r, b, V, col);
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
