// K2 / K2' / K3a — see bcr.cuh for the algorithm and the reference call sites replaced.
#include "bcr.cuh"

#include <algorithm>

namespace lgpu {

// ============================================================================= plan
BcrPlan make_bcr_plan(int n, int first_stage_m, int next_stage_m, int top_max_rows) {
  BcrPlan p;
  p.n = n;
  size_t f = 0, w = 0;
  int active = n;
  while (active >= 2) {
    BcrLevel lv{};
    lv.n_active = active;
    lv.n_elim = active / 2;
    lv.n_kept = active - lv.n_elim;
    lv.off_dinv = f; f += lv.n_elim;
    lv.off_lkuk = f; f += 2 * static_cast<size_t>(lv.n_kept);
    lv.off_glgu = f; f += 2 * static_cast<size_t>(lv.n_elim);
    lv.off_work = w;
    if (!p.levels.empty()) w += 3 * static_cast<size_t>(active);   // level 0 reads A, B directly
    p.levels.push_back(lv);
    active = lv.n_kept;
  }
  p.off_root = f; f += 1;
  p.factor_blocks = f;
  p.work_blocks = w + 3;   // + the root row
  // stages
  const int nl = static_cast<int>(p.levels.size());
  int l0 = 0;
  size_t rhs = 0, delta = 0;
  bool first = true;
  while (true) {
    const int n0 = l0 < nl ? p.levels[l0].n_active : 1;
    BcrStage st{};
    st.l0 = l0;
    st.n0 = n0;
    st.off_rin = rhs;
    st.off_delta = delta;
    if (n0 <= top_max_rows || l0 >= nl) {
      st.m = nl - l0;        // all remaining levels; 2^m >= n0
      st.nchunks = 1;
      rhs += n0;
      p.stages.push_back(st);
      break;
    }
    st.m = std::min(first ? first_stage_m : next_stage_m, nl - l0);
    const int C = 1 << st.m;
    st.nchunks = (n0 - 1 + C - 1) / C;
    rhs += n0;
    delta += 2 * (static_cast<size_t>(st.nchunks) + 1);
    p.stages.push_back(st);
    l0 += st.m;
    first = false;
  }
  p.rhs_vecs = rhs;
  p.delta_vecs = delta + 2;
  return p;
}

// ============================================================================ device
namespace {

__device__ __forceinline__ cd ldg_cd(const cd* p) {
  const double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return cd{v.x, v.y};
}

// y = M x for one 16x16 column-major block; x in shared memory.  Lane l owns row l & 15 and
// the 8 columns of half l >> 4; every load instruction of the warp covers two full 256-byte
// columns.  All lanes return the finished row value.
__device__ __forceinline__ cd block_mv(const cd* __restrict__ M, const cd* xs, int lane) {
  const int i = lane & 15, h = lane >> 4;
  cd m[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) m[c] = ldg_cd(M + (h * 8 + c) * BLK + i);
  cd acc{0.0, 0.0};
#pragma unroll
  for (int c = 0; c < 8; ++c) cfma(acc, m[c], xs[h * 8 + c]);
  acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
  acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
  return acc;
}

// y = M0 x0 + M1 x1 with all 16 loads in flight together
__device__ __forceinline__ cd block_mv2(const cd* __restrict__ M0, const cd* x0,
                                        const cd* __restrict__ M1, const cd* x1, int lane) {
  const int i = lane & 15, h = lane >> 4;
  cd a[8], b[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) a[c] = ldg_cd(M0 + (h * 8 + c) * BLK + i);
#pragma unroll
  for (int c = 0; c < 8; ++c) b[c] = ldg_cd(M1 + (h * 8 + c) * BLK + i);
  cd acc{0.0, 0.0};
#pragma unroll
  for (int c = 0; c < 8; ++c) cfma(acc, a[c], x0[h * 8 + c]);
#pragma unroll
  for (int c = 0; c < 8; ++c) cfma(acc, b[c], x1[h * 8 + c]);
  acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
  acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
  return acc;
}

constexpr int MAX_STAGE_LEVELS = 12;

struct LevelRef {
  size_t off_dinv, off_lkuk, off_glgu;
};

struct StageArgs {
  int l0, m, n0, nchunks;
  LevelRef lv[MAX_STAGE_LEVELS];
  size_t off_root;
  const cd* factors;
  const cd* rin;        // compact rhs of this stage (n0 x 16)
  const cd* dprev_l;    // previous stage's deltas (indexed by this stage's positions) or null
  const cd* dprev_r;
  cd* rout;             // compact rhs of the next stage
  cd* dl;               // this stage's deltas
  cd* dr;
  cd* yvec;             // global (n x 16)
  cd* xvec;             // global (n x 16)
};

// Load the stage right-hand side of the chunk (positions base .. base + C) into shared memory.
__device__ __forceinline__ void load_chunk_rhs(const StageArgs& a, int base, int C, cd* r) {
  for (int e = threadIdx.x; e < (C + 1) * BLK; e += blockDim.x) {
    const int q = e >> 4, i = e & 15;
    const int P = base + q;
    cd v{0.0, 0.0};
    if (P < a.n0) {
      v = a.rin[static_cast<size_t>(P) * BLK + i];
      if (a.dprev_l) {
        v -= a.dprev_l[static_cast<size_t>(P) * BLK + i];
        if (P > 0) v -= a.dprev_r[static_cast<size_t>(P - 1) * BLK + i];
      }
    }
    r[e] = v;
  }
}

// Forward elimination of one chunk over the stage's m levels.  TOP: the chunk is the whole
// remaining system (no right separator, the left "separator" row 0 is an ordinary kept row).
template <bool TOP>
__device__ __forceinline__ void chunk_forward(const StageArgs& a, int base, int C, cd* r, cd* y,
                                              cd* dl, cd* dr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int lam = 0; lam < a.m; ++lam) {
    const int s = 1 << lam;
    const int ne = C >> (lam + 1);
    const LevelRef lv = a.lv[lam];
    // (a) eliminated rows: y_q = Dinv_q r_q
    for (int t = warp; t < ne; t += nwarps) {
      const int q = s * (2 * t + 1);
      const int P = base + q;
      if (P >= a.n0) continue;
      const size_t te = static_cast<size_t>(P >> lam) >> 1;
      const cd v = block_mv(a.factors + (lv.off_dinv + te) * BLK2, r + q * BLK, lane);
      if (lane < BLK) {
        y[q * BLK + lane] = v;
        a.yvec[(static_cast<size_t>(P) << a.l0) * BLK + lane] = v;
      }
    }
    __syncthreads();
    // (b) kept rows: r_q -= Lk_q y_{q-s} + Uk_q y_{q+s}
    for (int u = warp; u <= ne; u += nwarps) {
      const int q = 2 * s * u;
      const int P = base + q;
      if (P >= a.n0) continue;
      const size_t ue = static_cast<size_t>(P >> lam) >> 1;
      const cd* Lk = a.factors + (lv.off_lkuk + 2 * ue) * BLK2;
      const cd* Uk = Lk + BLK2;
      const bool has_l = q > 0;
      const bool has_r = q < C && P + s < a.n0;
      cd acc{0.0, 0.0};
      if (has_l && has_r) acc = block_mv2(Lk, y + (q - s) * BLK, Uk, y + (q + s) * BLK, lane);
      else if (has_l) acc = block_mv(Lk, y + (q - s) * BLK, lane);
      else if (has_r) acc = block_mv(Uk, y + (q + s) * BLK, lane);
      if (lane < BLK) {
        if (!TOP && q == 0) dl[lane] += acc;
        else if (!TOP && q == C) dr[lane] += acc;
        else r[q * BLK + lane] -= acc;
      }
    }
    __syncthreads();
  }
}

// Back substitution of one chunk: x_q = y_q - GL_q x_{q-s} - GU_q x_{q+s}
__device__ __forceinline__ void chunk_backward(const StageArgs& a, int base, int C, cd* xs,
                                               const cd* ys /*smem y or null -> yvec*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int lam = a.m - 1; lam >= 0; --lam) {
    const int s = 1 << lam;
    const int ne = C >> (lam + 1);
    const LevelRef lv = a.lv[lam];
    for (int t = warp; t < ne; t += nwarps) {
      const int q = s * (2 * t + 1);
      const int P = base + q;
      if (P >= a.n0) continue;
      const size_t te = static_cast<size_t>(P >> lam) >> 1;
      const cd* GL = a.factors + (lv.off_glgu + 2 * te) * BLK2;
      const bool has_r = P + s < a.n0;   // q + s <= C always
      const size_t g = (static_cast<size_t>(P) << a.l0) * BLK;
      cd yv{0.0, 0.0};
      if (lane < BLK) yv = ys ? ys[q * BLK + lane] : a.yvec[g + lane];
      cd acc;
      if (has_r) acc = block_mv2(GL, xs + (q - s) * BLK, GL + BLK2, xs + (q + s) * BLK, lane);
      else acc = block_mv(GL, xs + (q - s) * BLK, lane);
      if (lane < BLK) {
        const cd xv = yv - acc;
        xs[q * BLK + lane] = xv;
        a.xvec[g + lane] = xv;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) bcr_fwd_stage_kernel(StageArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = 1 << a.m;
  cd* r = reinterpret_cast<cd*>(smem_raw);
  cd* y = r + (C + 1) * BLK;
  cd* dl = y + (C + 1) * BLK;
  cd* dr = dl + BLK;
  const int k = blockIdx.x, base = k * C;
  load_chunk_rhs(a, base, C, r);
  if (threadIdx.x < 2 * BLK) dl[threadIdx.x] = cd{0.0, 0.0};   // dl and dr are contiguous
  __syncthreads();
  if (threadIdx.x < BLK) {
    a.rout[static_cast<size_t>(k) * BLK + threadIdx.x] = r[threadIdx.x];
    if (k == a.nchunks - 1 && base + C < a.n0)
      a.rout[static_cast<size_t>(k + 1) * BLK + threadIdx.x] = r[C * BLK + threadIdx.x];
  }
  chunk_forward<false>(a, base, C, r, y, dl, dr);
  if (threadIdx.x < BLK) {
    a.dl[static_cast<size_t>(k) * BLK + threadIdx.x] = dl[threadIdx.x];
    a.dr[static_cast<size_t>(k) * BLK + threadIdx.x] = dr[threadIdx.x];
  }
}

// Single CTA: forward over all remaining levels, root solve, back substitution.
__global__ void __launch_bounds__(256) bcr_top_stage_kernel(StageArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = 1 << a.m;
  cd* r = reinterpret_cast<cd*>(smem_raw);
  cd* y = r + (C + 1) * BLK;
  load_chunk_rhs(a, 0, C, r);
  __syncthreads();
  chunk_forward<true>(a, 0, C, r, y, nullptr, nullptr);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    const cd v = block_mv(a.factors + a.off_root * BLK2, r, lane);
    if (lane < BLK) a.xvec[lane] = v;
    __syncwarp();
    if (lane < BLK) r[lane] = v;   // reuse r as the solution buffer (row 0)
  }
  __syncthreads();
  chunk_backward(a, 0, C, r, y);
}

__global__ void __launch_bounds__(256) bcr_bwd_stage_kernel(StageArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = 1 << a.m;
  cd* xs = reinterpret_cast<cd*>(smem_raw);
  const int k = blockIdx.x, base = k * C;
  if (threadIdx.x < BLK) {
    xs[threadIdx.x] = a.xvec[(static_cast<size_t>(base) << a.l0) * BLK + threadIdx.x];
  } else if (threadIdx.x < 2 * BLK) {
    const int i = threadIdx.x - BLK;
    cd v{0.0, 0.0};
    if (base + C < a.n0) v = a.xvec[(static_cast<size_t>(base + C) << a.l0) * BLK + i];
    xs[C * BLK + i] = v;
  }
  __syncthreads();
  chunk_backward(a, base, C, xs, nullptr);
}

// ------------------------------------------------------------------ factorisation kernels
struct FactorArgs {
  const cd* A;
  const cd* B;
  cd sigma;
  const cd* src;       // (L, D, U) rows of this level, or null at level 0 (use A - sigma*B)
  cd* dst;             // (L, D, U) rows of the next level
  cd* factors;
  size_t off_dinv, off_lkuk, off_glgu;
  int n_active;
  int row_stride;      // 2^level, for the singular-pivot report
  int32_t* info;
};

__device__ __forceinline__ cd load_entry(const FactorArgs& a, size_t row, int e) {
  if (a.src) return a.src[row * 3 * BLK2 + e];
  const size_t off = row * 3 * BLK2 + e;
  return a.A[off] - a.sigma * a.B[off];
}

constexpr int GJ_LD = 65;   // padded row length of the augmented matrix

// Pivoted Gauss-Jordan on [D | L | U | I] (16 x 64) of one eliminated row:
// -> [I | D^-1 L | D^-1 U | D^-1].  ROOT: only D^-1 is wanted.
template <bool ROOT>
__global__ void __launch_bounds__(256) bcr_factor_elim_kernel(FactorArgs a) {
  __shared__ cd W[BLK * GJ_LD];
  __shared__ cd fcol[BLK];
  __shared__ int piv_row;
  const int tid = threadIdx.x;
  const size_t te = blockIdx.x;
  const size_t row = ROOT ? 0 : 2 * te + 1;
  for (int e = tid; e < 3 * BLK2; e += 256) {
    const int blk = e >> 8, c = (e >> 4) & 15, i = e & 15;
    const int coff = blk == 1 ? 0 : (blk == 0 ? 16 : 32);   // storage order: sub, diag, super
    W[i * GJ_LD + coff + c] = load_entry(a, row, e);
  }
  {
    const int i = tid >> 4, c = tid & 15;
    W[i * GJ_LD + 48 + c] = cd{i == c ? 1.0 : 0.0, 0.0};
  }
  __syncthreads();
  for (int k = 0; k < BLK; ++k) {
    if (tid < 32) {
      // partial pivoting: largest |W[i][k]| over i >= k
      double best = (tid >= k && tid < BLK) ? abs2(W[tid * GJ_LD + k]) : -1.0;
      int bi = tid;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (tid == 0) {
        piv_row = bi;
        if (!(best > 0.0)) {   // exactly singular (or NaN): report like zgbtrf info > 0, keep going
          atomicCAS(a.info, 0, static_cast<int>(row) * a.row_stride + 1);
          W[bi * GJ_LD + k] = cd{2.2250738585072014e-308, 0.0};
        }
      }
    }
    __syncthreads();
    const int p = piv_row;
    if (p != k && tid < 64) {
      const cd t0 = W[k * GJ_LD + tid];
      W[k * GJ_LD + tid] = W[p * GJ_LD + tid];
      W[p * GJ_LD + tid] = t0;
    }
    __syncthreads();
    const cd pinv = crecip(W[k * GJ_LD + k]);
    if (tid >= 64 && tid < 64 + BLK) fcol[tid - 64] = W[(tid - 64) * GJ_LD + k];
    __syncthreads();
    if (tid < 64) W[k * GJ_LD + tid] = W[k * GJ_LD + tid] * pinv;
    __syncthreads();
    for (int e = tid; e < BLK * 64; e += 256) {
      const int i = e >> 6, c = e & 63;
      if (i != k) cfms(W[i * GJ_LD + c], fcol[i], W[k * GJ_LD + c]);
    }
    __syncthreads();
  }
  const int i = tid & 15, c = tid >> 4;
  if (ROOT) {
    a.factors[a.off_dinv * BLK2 + c * BLK + i] = W[i * GJ_LD + 48 + c];
  } else {
    a.factors[(a.off_dinv + te) * BLK2 + c * BLK + i] = W[i * GJ_LD + 48 + c];
    a.factors[(a.off_glgu + 2 * te) * BLK2 + c * BLK + i] = W[i * GJ_LD + 16 + c];
    a.factors[(a.off_glgu + 2 * te + 1) * BLK2 + c * BLK + i] = W[i * GJ_LD + 32 + c];
  }
}

// Schur-complement update of one kept row (position 2u of the level):
//   D' = D - L GU_left - U GL_right ;  L' = -L GL_left ;  U' = -U GU_right
__global__ void __launch_bounds__(256) bcr_factor_schur_kernel(FactorArgs a) {
  __shared__ cd S[7 * BLK2];   // L, D, U, GLl, GUl, GLr, GUr
  const int tid = threadIdx.x;
  const size_t u = blockIdx.x;
  const size_t row = 2 * u;
  const bool has_l = u >= 1;
  const bool has_r = 2 * u + 1 < static_cast<size_t>(a.n_active);
  for (int e = tid; e < 3 * BLK2; e += 256) S[e] = load_entry(a, row, e);
  {
    const cd z{0.0, 0.0};
    const cd* gl = a.factors + (a.off_glgu + 2 * (u - 1)) * BLK2;
    const cd* gr = a.factors + (a.off_glgu + 2 * u) * BLK2;
    S[3 * BLK2 + tid] = has_l ? gl[tid] : z;
    S[4 * BLK2 + tid] = has_l ? gl[BLK2 + tid] : z;
    S[5 * BLK2 + tid] = has_r ? gr[tid] : z;
    S[6 * BLK2 + tid] = has_r ? gr[BLK2 + tid] : z;
  }
  __syncthreads();
  // the pre-update couplings drive the forward sweep of every solve
  a.factors[(a.off_lkuk + 2 * u) * BLK2 + tid] = S[tid];
  a.factors[(a.off_lkuk + 2 * u + 1) * BLK2 + tid] = S[2 * BLK2 + tid];
  const int i = tid & 15, c = tid >> 4;
  cd dn = S[BLK2 + tid], ln{0.0, 0.0}, un{0.0, 0.0};
#pragma unroll
  for (int k = 0; k < BLK; ++k) {
    const cd l = S[k * BLK + i], uu = S[2 * BLK2 + k * BLK + i];
    cfms(dn, l, S[4 * BLK2 + c * BLK + k]);
    cfms(dn, uu, S[5 * BLK2 + c * BLK + k]);
    cfms(ln, l, S[3 * BLK2 + c * BLK + k]);
    cfms(un, uu, S[6 * BLK2 + c * BLK + k]);
  }
  cd* out = a.dst + u * 3 * BLK2;
  out[tid] = ln;
  out[BLK2 + tid] = dn;
  out[2 * BLK2 + tid] = un;
}

// ------------------------------------------------------------------------- block matvec
// y_b = aa * (A x)_b + ab * (B x)_b + z_b ; one warp per block row, 8 rows per CTA.
template <bool USE_A, bool USE_B>
__global__ void __launch_bounds__(256)
block_matvec_kernel(int n, const cd* __restrict__ A, const cd* __restrict__ B, cd aa, cd ab,
                    const cd* __restrict__ x, const cd* __restrict__ z, cd* __restrict__ y) {
  __shared__ cd xs[8][3 * BLK];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * 8 + warp;
  if (b >= n) return;
  cd* xw = xs[warp];
  for (int e = lane; e < 3 * BLK; e += 32) {
    const int bb = b - 1 + (e >> 4);
    xw[e] = (bb >= 0 && bb < n) ? x[static_cast<size_t>(bb) * BLK + (e & 15)] : cd{0.0, 0.0};
  }
  __syncwarp();
  const int i = lane & 15, h = lane >> 4;
  cd acc_a{0.0, 0.0}, acc_b{0.0, 0.0};
  const size_t rowoff = static_cast<size_t>(b) * 3 * BLK2;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    if (USE_A) {
      cd m[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) m[c] = ldg_cd(A + rowoff + t * BLK2 + (h * 8 + c) * BLK + i);
#pragma unroll
      for (int c = 0; c < 8; ++c) cfma(acc_a, m[c], xw[t * BLK + h * 8 + c]);
    }
    if (USE_B) {
      cd m[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) m[c] = ldg_cd(B + rowoff + t * BLK2 + (h * 8 + c) * BLK + i);
#pragma unroll
      for (int c = 0; c < 8; ++c) cfma(acc_b, m[c], xw[t * BLK + h * 8 + c]);
    }
  }
  cd acc{0.0, 0.0};
  if (USE_A) acc += aa * acc_a;
  if (USE_B) acc += ab * acc_b;
  acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
  acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
  if (lane < BLK) {
    if (z) acc += z[static_cast<size_t>(b) * BLK + lane];
    y[static_cast<size_t>(b) * BLK + lane] = acc;
  }
}

StageArgs make_stage_args(const BcrPlan& plan, const BcrDevice& d, int s, const cd* b, cd* x) {
  const BcrStage& st = plan.stages[s];
  StageArgs a{};
  a.l0 = st.l0; a.m = st.m; a.n0 = st.n0; a.nchunks = st.nchunks;
  for (int lam = 0; lam < st.m; ++lam) {
    const BcrLevel& lv = plan.levels[st.l0 + lam];
    a.lv[lam] = LevelRef{lv.off_dinv, lv.off_lkuk, lv.off_glgu};
  }
  a.off_root = plan.off_root;
  a.factors = d.factors;
  a.rin = s == 0 ? b : d.rhs + st.off_rin * BLK;
  if (s > 0) {
    const BcrStage& pv = plan.stages[s - 1];
    a.dprev_l = d.delta + pv.off_delta * BLK;
    a.dprev_r = a.dprev_l + (static_cast<size_t>(pv.nchunks) + 1) * BLK;
  }
  const bool top = s == static_cast<int>(plan.stages.size()) - 1;
  if (!top) {
    a.rout = d.rhs + plan.stages[s + 1].off_rin * BLK;
    a.dl = d.delta + st.off_delta * BLK;
    a.dr = a.dl + (static_cast<size_t>(st.nchunks) + 1) * BLK;
  }
  a.yvec = d.yvec;
  a.xvec = x;
  return a;
}

}  // namespace

void bcr_factorize(const BcrPlan& plan, const BcrDevice& d, cd sigma, cudaStream_t stream,
                   int64_t* launches) {
  CUDA_CHECK(cudaMemsetAsync(d.info, 0, sizeof(int32_t), stream));
  const int nl = static_cast<int>(plan.levels.size());
  FactorArgs a{};
  a.A = d.A; a.B = d.B; a.sigma = sigma; a.factors = d.factors; a.info = d.info;
  for (int l = 0; l < nl; ++l) {
    const BcrLevel& lv = plan.levels[l];
    a.src = l == 0 ? nullptr : d.work + lv.off_work * BLK2;
    const size_t next_off = l + 1 < nl ? plan.levels[l + 1].off_work : plan.work_blocks - 3;
    a.dst = d.work + next_off * BLK2;
    a.off_dinv = lv.off_dinv; a.off_lkuk = lv.off_lkuk; a.off_glgu = lv.off_glgu;
    a.n_active = lv.n_active;
    a.row_stride = 1 << l;
    bcr_factor_elim_kernel<false><<<lv.n_elim, 256, 0, stream>>>(a);
    bcr_factor_schur_kernel<<<lv.n_kept, 256, 0, stream>>>(a);
    *launches += 2;
  }
  // root: invert the last remaining diagonal block
  a.src = nl == 0 ? nullptr : d.work + (plan.work_blocks - 3) * BLK2;
  a.off_dinv = plan.off_root;
  a.n_active = 1;
  a.row_stride = 1;
  bcr_factor_elim_kernel<true><<<1, 256, 0, stream>>>(a);
  *launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

void bcr_solve(const BcrPlan& plan, const BcrDevice& d, const cd* b, cd* x, cudaStream_t stream,
               int64_t* launches) {
  const int ns = static_cast<int>(plan.stages.size());
  for (int s = 0; s < ns - 1; ++s) {
    const StageArgs a = make_stage_args(plan, d, s, b, x);
    const int C = 1 << a.m;
    const size_t smem = sizeof(cd) * (2 * (C + 1) * BLK + 2 * BLK);
    bcr_fwd_stage_kernel<<<a.nchunks, 256, smem, stream>>>(a);
  }
  {
    const StageArgs a = make_stage_args(plan, d, ns - 1, b, x);
    const int C = 1 << a.m;
    const size_t smem = sizeof(cd) * (2 * (C + 1) * BLK);
    bcr_top_stage_kernel<<<1, 256, smem, stream>>>(a);
  }
  for (int s = ns - 2; s >= 0; --s) {
    const StageArgs a = make_stage_args(plan, d, s, b, x);
    const int C = 1 << a.m;
    const size_t smem = sizeof(cd) * ((C + 1) * BLK);
    bcr_bwd_stage_kernel<<<a.nchunks, 256, smem, stream>>>(a);
  }
  *launches += 2 * (ns - 1) + 1;
  CUDA_CHECK(cudaGetLastError());
}

void block_matvec(int n, const cd* A, const cd* B, cd aa, cd ab, const cd* x, const cd* z, cd* y,
                  cudaStream_t stream, int64_t* launches) {
  const int ctas = (n + 7) / 8;
  const bool use_a = aa.x != 0.0 || aa.y != 0.0, use_b = ab.x != 0.0 || ab.y != 0.0;
  if (use_a && use_b) block_matvec_kernel<true, true><<<ctas, 256, 0, stream>>>(n, A, B, aa, ab, x, z, y);
  else if (use_a) block_matvec_kernel<true, false><<<ctas, 256, 0, stream>>>(n, A, B, aa, ab, x, z, y);
  else block_matvec_kernel<false, true><<<ctas, 256, 0, stream>>>(n, A, B, aa, ab, x, z, y);
  *launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace lgpu
