// Microbenchmark: latency of the unit-diagonal upper triangular solve of one 32 x 32 (and 64 x 64) complex block by ONE
// warp, the step the upper solve stages chain 9 + 1 times per operator application (slu.cu).  Variants are timed as a
// dependent chain of solves (cycles per solve, clock64) and compared bit by bit with the column-at-a-time substitution.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o trisolve trisolve.cu && ./trisolve
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

struct cd { double x, y; };
__host__ __device__ inline cd operator*(cd a, cd b) { return cd{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__host__ __device__ inline void cfms(cd& a, cd b, cd c) {
  a.x = fma(-b.x, c.x, a.x); a.x = fma(b.y, c.y, a.x);
  a.y = fma(-b.x, c.y, a.y); a.y = fma(-b.y, c.x, a.y);
}
constexpr int SB = 32;
constexpr int TRI = SB * (SB + 1) / 2;
__host__ __device__ inline int tri_up_off(int k) { return (k * (k + 1)) / 2; }
__device__ __forceinline__ cd shfl_cd(cd v, int src) {
  return cd{__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)};
}

// V1: one column per step
__device__ __forceinline__ cd solve_seq(const cd* U, cd r, int lane) {
  r = r * U[tri_up_off(lane) + lane];
  cd u[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) u[j] = U[tri_up_off(SB - 1 - j) + min(lane, SB - 1 - j)];
#pragma unroll
  for (int k = SB - 1; k >= 1; --k) {
    const cd xk = shfl_cd(r, k);
    const cd uk = u[(SB - 1 - k) & 3];
    if (k >= 5) u[(SB - 1 - k) & 3] = U[tri_up_off(k - 4) + min(lane, k - 4)];
    if (lane < k) cfms(r, uk, xk);
  }
  return r;
}

// V0: blocks of four columns, as in slu.cu before this study
__device__ __forceinline__ cd solve_blk4(const cd* U, cd r, int lane) {
  r = r * U[tri_up_off(lane) + lane];
#pragma unroll
  for (int kb = SB / 4 - 1; kb >= 0; --kb) {
    const int c0 = 4 * kb;
    const cd* k1 = U + tri_up_off(c0 + 1);
    const cd* k2 = U + tri_up_off(c0 + 2);
    const cd* k3 = U + tri_up_off(c0 + 3);
    const int li = min(lane, c0);
    const cd w0 = U[tri_up_off(c0) + li], w1 = k1[li], w2 = k2[li], w3 = k3[li];
    const cd u01 = k1[c0], u02 = k2[c0], u12 = k2[c0 + 1], u03 = k3[c0], u13 = k3[c0 + 1], u23 = k3[c0 + 2];
    const cd a0 = shfl_cd(r, c0), a1 = shfl_cd(r, c0 + 1), a2 = shfl_cd(r, c0 + 2), x3 = shfl_cd(r, c0 + 3);
    cd x2 = a2; cfms(x2, u23, x3);
    cd x1 = a1; cfms(x1, u13, x3); cfms(x1, u12, x2);
    cd x0 = a0; cfms(x0, u03, x3); cfms(x0, u02, x2); cfms(x0, u01, x1);
    if (lane < c0) {
      cfms(r, w3, x3); cfms(r, w2, x2); cfms(r, w1, x1); cfms(r, w0, x0);
    } else if (lane < c0 + 4) {
      const int j = lane - c0;
      r = j == 0 ? x0 : (j == 1 ? x1 : (j == 2 ? x2 : x3));
    }
  }
  return r;
}

struct Blk { cd w0, w1, w2, w3, u01, u02, u12, u03, u13, u23; };
__device__ __forceinline__ Blk load_blk(const cd* __restrict__ U, int c0, int lane) {
  const cd* k0 = U + tri_up_off(c0);
  const cd* k1 = U + tri_up_off(c0 + 1);
  const cd* k2 = U + tri_up_off(c0 + 2);
  const cd* k3 = U + tri_up_off(c0 + 3);
  const int li = min(lane, c0);
  const bool above = lane < c0;
  const cd z{0.0, 0.0};
  Blk b;
  b.w0 = above ? k0[li] : z; b.w1 = above ? k1[li] : z; b.w2 = above ? k2[li] : z; b.w3 = above ? k3[li] : z;
  b.u01 = k1[c0]; b.u02 = k2[c0]; b.u12 = k2[c0 + 1]; b.u03 = k3[c0]; b.u13 = k3[c0 + 1]; b.u23 = k3[c0 + 2];
  return b;
}
__device__ __forceinline__ cd step_blk(const Blk& b, cd r, int c0, int lane) {
  const cd a0 = shfl_cd(r, c0), a1 = shfl_cd(r, c0 + 1), a2 = shfl_cd(r, c0 + 2), x3 = shfl_cd(r, c0 + 3);
  cfms(r, b.w3, x3);
  cd x2 = a2; cfms(x2, b.u23, x3);
  cfms(r, b.w2, x2);
  cd x1 = a1; cfms(x1, b.u13, x3); cfms(x1, b.u12, x2);
  cfms(r, b.w1, x1);
  cd x0 = a0; cfms(x0, b.u03, x3); cfms(x0, b.u02, x2); cfms(x0, b.u01, x1);
  cfms(r, b.w0, x0);
  const int j = lane - c0;
  r = j == 0 ? x0 : r; r = j == 1 ? x1 : r; r = j == 2 ? x2 : r; r = j == 3 ? x3 : r;
  return r;
}
// V2: same arithmetic per entry (bit-identical), no divergent branch (rows at / below the block multiply zeros),
// next block's entries loaded before the current block's chain
__device__ __forceinline__ cd solve_pipe(const cd* __restrict__ U, cd r, int lane) {
  r = r * U[tri_up_off(lane) + lane];
  Blk cur = load_blk(U, SB - 4, lane);
#pragma unroll
  for (int kb = SB / 4 - 1; kb >= 0; --kb) {
    Blk nxt = cur;
    if (kb > 0) nxt = load_blk(U, 4 * kb - 4, lane);
    r = step_blk(cur, r, 4 * kb, lane);
    cur = nxt;
  }
  return r;
}
// V3: every block's entries in registers before the right-hand side is touched (the factor is static: in the solve
// kernels this load can happen while the warp waits for its inputs)
struct AllBlk { Blk b[SB / 4]; };
__device__ __forceinline__ void load_all(const cd* __restrict__ U, int lane, AllBlk& a) {
#pragma unroll
  for (int kb = 0; kb < SB / 4; ++kb) a.b[kb] = load_blk(U, 4 * kb, lane);
}
__device__ __forceinline__ cd solve_regs(const AllBlk& a, cd dinv, cd r, int lane) {
  r = r * dinv;
#pragma unroll
  for (int kb = SB / 4 - 1; kb >= 0; --kb) r = step_blk(a.b[kb], r, 4 * kb, lane);
  return r;
}

// V4: blocks of NB columns without selects: lane i multiplies column c by U(i, c) if i < c and by zero otherwise, so the
// lanes inside the block obtain their own unknown from the same update as the rows above it
template <int NB>
struct MBlk { cd w[NB]; cd u[NB][NB]; };   // u[j][m], j < m: U(c0 + j, c0 + m)
template <int NB>
__device__ __forceinline__ MBlk<NB> load_mblk(const cd* __restrict__ U, int c0, int lane) {
  MBlk<NB> b;
  const cd z{0.0, 0.0};
#pragma unroll
  for (int m = 0; m < NB; ++m) {
    const cd* col = U + tri_up_off(c0 + m);
    const cd v = col[min(lane, c0 + m)];
    b.w[m] = lane < c0 + m ? v : z;
#pragma unroll
    for (int j = 0; j < m; ++j) b.u[j][m] = col[c0 + j];
  }
  return b;
}
template <int NB>
__device__ __forceinline__ cd step_mblk(const MBlk<NB>& b, cd r, int c0) {
  cd x[NB];
#pragma unroll
  for (int m = 0; m < NB; ++m) x[m] = shfl_cd(r, c0 + m);
#pragma unroll
  for (int m = NB - 1; m >= 0; --m) {
    // x[m] is final here
    cfms(r, b.w[m], x[m]);
#pragma unroll
    for (int j = m - 1; j >= 0; --j) cfms(x[j], b.u[j][m], x[m]);
  }
  return r;
}
template <int NB>
__device__ __forceinline__ cd solve_masked(const cd* __restrict__ U, cd r, int lane) {
  r = r * U[tri_up_off(lane) + lane];
  MBlk<NB> cur = load_mblk<NB>(U, SB - NB, lane);
#pragma unroll
  for (int kb = SB / NB - 1; kb >= 0; --kb) {
    MBlk<NB> nxt = cur;
    if (kb > 0) nxt = load_mblk<NB>(U, NB * kb - NB, lane);
    r = step_mblk<NB>(cur, r, NB * kb);
    cur = nxt;
  }
  return r;
}

template <int V>
__global__ void __launch_bounds__(32, 1) bench32(const cd* Ug, const cd* rg, cd* out, long long* cyc, int iters) {
  __shared__ cd U[TRI];
  const int lane = threadIdx.x;
  for (int e = lane; e < TRI; e += 32) U[e] = Ug[e];
  __syncwarp();
  cd r = rg[lane];
  AllBlk all;
  cd dinv{0.0, 0.0};
  if (V == 3) { load_all(U, lane, all); dinv = U[tri_up_off(lane) + lane]; }
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (V == 0) r = solve_blk4(U, r, lane);
    if (V == 1) r = solve_seq(U, r, lane);
    if (V == 2) r = solve_pipe(U, r, lane);
    if (V == 3) r = solve_regs(all, dinv, r, lane);
    if (V == 4) r = solve_masked<4>(U, r, lane);
    if (V == 5) r = solve_masked<2>(U, r, lane);
    if (V == 6) r = solve_masked<8>(U, r, lane);
  }
  const long long t1 = clock64();
  out[lane] = r;
  if (lane == 0) *cyc = t1 - t0;
}

// ---- 64 x 64, column-major with leading dimension 64, rows lane and lane + 32
__device__ __forceinline__ void solve64_old(const cd* U, cd& y0, cd& y1, int lane) {
  y0 = y0 * U[lane * 64 + lane];
  y1 = y1 * U[(lane + 32) * 64 + lane + 32];
#pragma unroll 4
  for (int kb = 15; kb >= 0; --kb) {
    const int c0 = 4 * kb;
    const bool hi = kb >= 8;
    const cd* k0 = U + c0 * 64;
    const cd* k1 = k0 + 64;
    const cd* k2 = k0 + 128;
    const cd* k3 = k0 + 192;
    const cd u01 = k1[c0], u02 = k2[c0], u12 = k2[c0 + 1], u03 = k3[c0], u13 = k3[c0 + 1], u23 = k3[c0 + 2];
    const cd src = hi ? y1 : y0;
    const cd a0 = shfl_cd(src, c0 & 31), a1 = shfl_cd(src, (c0 + 1) & 31), a2 = shfl_cd(src, (c0 + 2) & 31),
             x3 = shfl_cd(src, (c0 + 3) & 31);
    cd x2 = a2; cfms(x2, u23, x3);
    cd x1 = a1; cfms(x1, u13, x3); cfms(x1, u12, x2);
    cd x0 = a0; cfms(x0, u03, x3); cfms(x0, u02, x2); cfms(x0, u01, x1);
    if (hi || lane < c0) {
      cfms(y0, k3[lane], x3); cfms(y0, k2[lane], x2); cfms(y0, k1[lane], x1); cfms(y0, k0[lane], x0);
    } else if (lane < c0 + 4) {
      const int j = lane - c0;
      y0 = j == 0 ? x0 : (j == 1 ? x1 : (j == 2 ? x2 : x3));
    }
    if (hi) {
      const int i = lane + 32;
      if (i < c0) {
        cfms(y1, k3[i], x3); cfms(y1, k2[i], x2); cfms(y1, k1[i], x1); cfms(y1, k0[i], x0);
      } else if (i < c0 + 4) {
        const int j = i - c0;
        y1 = j == 0 ? x0 : (j == 1 ? x1 : (j == 2 ? x2 : x3));
      }
    }
  }
}
struct Blk64 { cd p0, p1, p2, p3, q0, q1, q2, q3, u01, u02, u12, u03, u13, u23; };   // p: row lane, q: row lane + 32
template <bool HI>
__device__ __forceinline__ Blk64 load_blk64(const cd* __restrict__ U, int c0, int lane) {
  const cd* k0 = U + c0 * 64;
  const cd z{0.0, 0.0};
  Blk64 b;
  const bool pa = lane < c0;   // row `lane` is above the block
  b.p0 = pa ? k0[lane] : z; b.p1 = pa ? k0[64 + lane] : z; b.p2 = pa ? k0[128 + lane] : z; b.p3 = pa ? k0[192 + lane] : z;
  if (HI) {
    const bool qa = lane + 32 < c0;
    const int i = min(lane + 32, c0);
    b.q0 = qa ? k0[i] : z; b.q1 = qa ? k0[64 + i] : z; b.q2 = qa ? k0[128 + i] : z; b.q3 = qa ? k0[192 + i] : z;
  }
  b.u01 = k0[64 + c0]; b.u02 = k0[128 + c0]; b.u12 = k0[128 + c0 + 1];
  b.u03 = k0[192 + c0]; b.u13 = k0[192 + c0 + 1]; b.u23 = k0[192 + c0 + 2];
  return b;
}
template <bool HI>
__device__ __forceinline__ void step_blk64(const Blk64& b, cd& y0, cd& y1, int c0, int lane) {
  const cd src = HI ? y1 : y0;
  const cd a0 = shfl_cd(src, c0 & 31), a1 = shfl_cd(src, (c0 + 1) & 31), a2 = shfl_cd(src, (c0 + 2) & 31),
           x3 = shfl_cd(src, (c0 + 3) & 31);
  cfms(y0, b.p3, x3); if (HI) cfms(y1, b.q3, x3);
  cd x2 = a2; cfms(x2, b.u23, x3);
  cfms(y0, b.p2, x2); if (HI) cfms(y1, b.q2, x2);
  cd x1 = a1; cfms(x1, b.u13, x3); cfms(x1, b.u12, x2);
  cfms(y0, b.p1, x1); if (HI) cfms(y1, b.q1, x1);
  cd x0 = a0; cfms(x0, b.u03, x3); cfms(x0, b.u02, x2); cfms(x0, b.u01, x1);
  cfms(y0, b.p0, x0); if (HI) cfms(y1, b.q0, x0);
  const int j = lane - (c0 & 31);
  cd& t = HI ? y1 : y0;
  t = j == 0 ? x0 : t; t = j == 1 ? x1 : t; t = j == 2 ? x2 : t; t = j == 3 ? x3 : t;
}
__device__ __forceinline__ void solve64_pipe(const cd* __restrict__ U, cd& y0, cd& y1, int lane) {
  y0 = y0 * U[lane * 64 + lane];
  y1 = y1 * U[(lane + 32) * 64 + lane + 32];
  Blk64 cur = load_blk64<true>(U, 60, lane);
#pragma unroll
  for (int kb = 15; kb >= 8; --kb) {
    Blk64 nxt = kb > 8 ? load_blk64<true>(U, 4 * kb - 4, lane) : load_blk64<false>(U, 4 * kb - 4, lane);
    step_blk64<true>(cur, y0, y1, 4 * kb, lane);
    cur = nxt;
  }
#pragma unroll
  for (int kb = 7; kb >= 0; --kb) {
    Blk64 nxt = cur;
    if (kb > 0) nxt = load_blk64<false>(U, 4 * kb - 4, lane);
    step_blk64<false>(cur, y0, y1, 4 * kb, lane);
    cur = nxt;
  }
}
template <int V>
__global__ void __launch_bounds__(32, 1) bench64(const cd* Ug, const cd* rg, cd* out, long long* cyc, int iters) {
  extern __shared__ cd U64[];
  const int lane = threadIdx.x;
  for (int e = lane; e < 64 * 64; e += 32) U64[e] = Ug[e];
  __syncwarp();
  cd y0 = rg[lane], y1 = rg[lane + 32];
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (V == 0) solve64_old(U64, y0, y1, lane);
    if (V == 1) solve64_pipe(U64, y0, y1, lane);
  }
  const long long t1 = clock64();
  out[lane] = y0; out[lane + 32] = y1;
  if (lane == 0) *cyc = t1 - t0;
}

int main() {
  cd *U, *r, *out, *U64;
  long long* cyc;
  cudaMallocManaged(&U, sizeof(cd) * TRI); cudaMallocManaged(&r, sizeof(cd) * 64); cudaMallocManaged(&out, sizeof(cd) * 64 * 8);
  cudaMallocManaged(&U64, sizeof(cd) * 64 * 64);
  cudaMallocManaged(&cyc, 8);
  srand(1);
  auto rnd = [] { return rand() / double(RAND_MAX) - 0.5; };
  // small off-diagonal entries and a unit "reciprocal diagonal": repeated solves stay bounded
  for (int c = 0; c < SB; ++c)
    for (int i = 0; i <= c; ++i) U[tri_up_off(c) + i] = i == c ? cd{1.0, 0.0} : cd{0.05 * rnd(), 0.05 * rnd()};
  for (int c = 0; c < 64; ++c)
    for (int i = 0; i < 64; ++i) U64[c * 64 + i] = i == c ? cd{1.0, 0.0} : (i < c ? cd{0.03 * rnd(), 0.03 * rnd()} : cd{0.0, 0.0});
  for (int i = 0; i < 64; ++i) r[i] = cd{rnd(), rnd()};
  const int iters = 200;
  const char* names[] = {"blocks of 4 (divergent, loads in the chain)", "column at a time", "blocks of 4, branch-free, pipelined loads",
                         "blocks of 4, factor in registers", "blocks of 4, masked coefficients (no selects)",
                         "blocks of 2, masked coefficients", "blocks of 8, masked coefficients"};
  cd ref[32];
  for (int v : {1, 0, 2, 3, 4, 5, 6}) {
    for (int rep = 0; rep < 2; ++rep) {
      if (v == 0) bench32<0><<<1, 32>>>(U, r, out, cyc, iters);
      if (v == 1) bench32<1><<<1, 32>>>(U, r, out, cyc, iters);
      if (v == 2) bench32<2><<<1, 32>>>(U, r, out, cyc, iters);
      if (v == 3) bench32<3><<<1, 32>>>(U, r, out, cyc, iters);
      if (v == 4) bench32<4><<<1, 32>>>(U, r, out, cyc, iters);
      if (v == 5) bench32<5><<<1, 32>>>(U, r, out, cyc, iters);
      if (v == 6) bench32<6><<<1, 32>>>(U, r, out, cyc, iters);
      cudaDeviceSynchronize();
    }
    bool same = true;
    if (v == 1) for (int i = 0; i < 32; ++i) ref[i] = out[i];
    for (int i = 0; i < 32; ++i) same = same && out[i].x == ref[i].x && out[i].y == ref[i].y;
    printf("32 x 32  %-48s %7.1f cycles per solve  %s  (%s)\n", names[v], double(*cyc) / iters, same ? "bit-identical" : "DIFFERENT",
           cudaGetErrorString(cudaGetLastError()));
  }
  cd ref64[64];
  cudaFuncSetAttribute(bench64<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 64 * 16);
  cudaFuncSetAttribute(bench64<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 64 * 16);
  for (int v : {0, 1}) {
    for (int rep = 0; rep < 2; ++rep) {
      if (v == 0) bench64<0><<<1, 32, 64 * 64 * 16>>>(U64, r, out, cyc, iters);
      if (v == 1) bench64<1><<<1, 32, 64 * 64 * 16>>>(U64, r, out, cyc, iters);
      cudaDeviceSynchronize();
    }
    bool same = true;
    if (v == 0) for (int i = 0; i < 64; ++i) ref64[i] = out[i];
    for (int i = 0; i < 64; ++i) same = same && out[i].x == ref64[i].x && out[i].y == ref64[i].y;
    printf("64 x 64  %-48s %7.1f cycles per solve  %s  (%s)\n", v == 0 ? "blocks of 4 (divergent)" : "blocks of 4, branch-free, pipelined loads",
           double(*cyc) / iters, same ? "bit-identical" : "DIFFERENT", cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
