// K3b / K3c — see arnoldi.cuh.
//
// All kernels here are pure streaming kernels over the N x ncv basis (HBM-bound).  Reductions
// are two-level (CTA partials, then the last CTA to finish sums the partials in a fixed
// order), so results are deterministic.
#include "arnoldi.cuh"

#include <algorithm>
#include <cstdlib>

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

// ---- staged Gram-Schmidt pass -----------------------------------------------------------------
// One kernel for the three passes of a CGS2 step:
//   <false, true>   h = V^H w                                   (first projection)
//   <true,  true>   w -= V h_in ; s = V^H w                     (correction + second projection)
//   <true,  false>  w -= V h_in ; ||w||                         (second correction + norm)
// A CTA owns a contiguous range of 64-row tiles (one CTA per SM, ranges differ by at most one
// tile).  The first ncols columns of a tile are one contiguous block of the basis, so a producer
// warp streams them with ONE cp.async.bulk per tile into a ring of shared-memory stages.  The
// 256 consumer threads are 64 rows x 4 column groups: a thread pulls its <= 16 entries of the
// tile into registers once and uses them for both the correction and the projection, so the
// fused middle pass reads V once; projections accumulate in registers over the whole tile range
// and are reduced across rows once per CTA.  Two-level fixed-order reductions (deterministic).
constexpr int PASS_THREADS = 288;   // warps 0-7 consume, warp 8 produces
constexpr int PASS_GROUPS = 4;      // column groups
constexpr int PASS_CPG = KRYLOV_PASS_MAXCOL / PASS_GROUPS;   // columns per thread
constexpr int PASS_T = KRYLOV_TILE;

template <bool UPDATE, bool DOTS>
__global__ void __launch_bounds__(PASS_THREADS, 1)
krylov_pass_kernel(BasisLayout L, const cd* __restrict__ V, int ncols, int nstages, cd* w, const cd* hin,
                   cd* __restrict__ partial, cd* hwork, cd* Hcol, int accumulate, double* scal,
                   unsigned int* ticket) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int sstride = (ncols + 1) * PASS_T;                             // stage: [ncols][64] of V, [64] of w
  cd* buf = reinterpret_cast<cd*>(smem_raw);                            // [nstages][sstride]
  cd* part = buf + static_cast<size_t>(nstages) * sstride;               // [2][4][64]
  cd* hs = part + 2 * PASS_GROUPS * PASS_T;                             // [KRYLOV_PASS_MAXCOL]
  uint64_t* full = reinterpret_cast<uint64_t*>(hs + KRYLOV_PASS_MAXCOL);
  uint64_t* empty = full + nstages;
  __shared__ double red[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t0 = static_cast<int>(static_cast<long long>(blockIdx.x) * L.ntiles / gridDim.x);
  const int t1 = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * L.ntiles / gridDim.x);
  if (tid == 0) {
    for (int i = 0; i < nstages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  if (UPDATE) for (int c = tid; c < ncols; c += PASS_THREADS) hs[c] = hin[c];
  __syncthreads();
  const int cpg = (ncols + PASS_GROUPS - 1) / PASS_GROUPS;
  const int r = tid & (PASS_T - 1), q = (tid >> 6) & 3;
  cd acc[PASS_CPG];
#pragma unroll
  for (int j = 0; j < PASS_CPG; ++j) acc[j] = cd{0.0, 0.0};
  double nrm = 0.0;
  if (warp == 8) {
    // ---- producer (one lane): tile t -> stage (t - t0) % nstages
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t vbytes = static_cast<uint32_t>(sizeof(cd) * ncols * PASS_T);
      for (int t = t0; t < t1; ++t) {
        if (t - t0 >= nstages) mbar_wait(&empty[stage], phase ^ 1u);
        cd* sb = buf + static_cast<size_t>(stage) * sstride;
        const uint32_t wbytes = static_cast<uint32_t>(sizeof(cd) * min(PASS_T, L.n - t * PASS_T));
        mbar_expect_tx(&full[stage], vbytes + wbytes);
        if (ncols > 0) bulk_g2s(sb, V + static_cast<size_t>(t) * L.ncv * PASS_T, vbytes, &full[stage]);
        bulk_g2s(sb + ncols * PASS_T, w + static_cast<size_t>(t) * PASS_T, wbytes, &full[stage]);
        if (++stage == nstages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ---- consumers
    int stage = 0;
    uint32_t phase = 0;
    int flip = 0;
    for (int t = t0; t < t1; ++t) {
      const int gi = t * PASS_T + r;
      const bool valid = gi < L.n;
      cd wi{0.0, 0.0};
      cd v[PASS_CPG];
      {
        mbar_wait(&full[stage], phase);
        const cd* sb = buf + static_cast<size_t>(stage) * sstride;
        if (valid) wi = sb[ncols * PASS_T + r];
#pragma unroll
        for (int j = 0; j < PASS_CPG; ++j) {
          const int c = q * cpg + j;
          v[j] = (j < cpg && c < ncols && valid) ? sb[c * PASS_T + r] : cd{0.0, 0.0};
        }
        mbar_release_slot(&empty[stage], lane);   // the tile now lives in registers
        if (++stage == nstages) { stage = 0; phase ^= 1u; }
      }
      if (UPDATE && ncols > 0) {
        cd p0{0.0, 0.0}, p1{0.0, 0.0};
#pragma unroll
        for (int j = 0; j < PASS_CPG; j += 2) {
          if (j < cpg) {
            const int c = q * cpg + j;
            cfma(p0, v[j], hs[min(c, ncols - 1)]);
            cfma(p1, v[j + 1], hs[min(c + 1, ncols - 1)]);
          }
        }
        cd* pp = part + flip * PASS_GROUPS * PASS_T;
        flip ^= 1;
        pp[q * PASS_T + r] = p0 + p1;
        asm volatile("bar.sync 1, 256;" ::: "memory");   // consumer warps only
        wi = wi - ((pp[r] + pp[PASS_T + r]) + (pp[2 * PASS_T + r] + pp[3 * PASS_T + r]));
        if (q == 0 && valid) w[gi] = wi;
      }
      if (DOTS) {
#pragma unroll
        for (int j = 0; j < PASS_CPG; ++j)
          if (j < cpg) cfmac(acc[j], v[j], wi);
      } else if (q == 0) {
        nrm += abs2(wi);
      }
    }
    if (DOTS) {
      // rows -> one value per (CTA, column): lanes, then the two warps of a column group
#pragma unroll
      for (int j = 0; j < PASS_CPG; ++j) {
        if (j < cpg) {
          acc[j].x = warp_sum(acc[j].x);
          acc[j].y = warp_sum(acc[j].y);
        }
      }
      cd* xw = part;   // [8 warps][PASS_CPG]
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < PASS_CPG; ++j) xw[warp * PASS_CPG + j] = acc[j];
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int e = tid; e < PASS_GROUPS * cpg; e += 256) {
        const int qq = e / cpg, j = e - qq * cpg;
        const int c = qq * cpg + j;
        if (c < ncols)
          partial[static_cast<size_t>(blockIdx.x) * PSTRIDE + c] = xw[(2 * qq) * PASS_CPG + j] + xw[(2 * qq + 1) * PASS_CPG + j];
      }
    } else {
      nrm = warp_sum(nrm);
      if (lane == 0) red[warp] = nrm;
    }
  }
  if (!DOTS) {
    __syncthreads();
    if (tid == 0) partial[static_cast<size_t>(blockIdx.x) * PSTRIDE + KRYLOV_MAXCOL] = cd{red[0] + red[1], 0.0};
  }
  if (last_block_done(ticket, true)) {
    if (DOTS) {
      // one warp per column: lanes stride over the CTA partials, fixed-order shuffle tree
      for (int c = warp; c < ncols; c += PASS_THREADS / 32) {
        cd s{0.0, 0.0};
        for (unsigned int b = lane; b < gridDim.x; b += 32) s += partial[static_cast<size_t>(b) * PSTRIDE + c];
        s.x = warp_sum(s.x);
        s.y = warp_sum(s.y);
        if (lane == 0) {
          hwork[c] = s;
          if (Hcol) Hcol[c] = accumulate ? Hcol[c] + s : s;
        }
      }
    } else if (warp == 0) {
      double s = 0.0;
      for (unsigned int b = lane; b < gridDim.x; b += 32)
        s += partial[static_cast<size_t>(b) * PSTRIDE + KRYLOV_MAXCOL].x;
      s = warp_sum(s);
      if (lane == 0) scal[0] = sqrt(s);
    }
    __syncthreads();
    if (tid == 0) *ticket = 0u;
  }
}

// ---- one cooperative launch per Arnoldi step ------------------------------------------------
// The three passes above and the normalisation of the new basis vector as ONE kernel (grid = one
// CTA per SM, co-resident): between the passes the CTAs meet at a device-wide barrier (a
// monotonic counter), every CTA sums the per-CTA partials in the same fixed order, and the
// producer warp keeps streaming the next pass's tiles while the consumers wait (V does not
// change during the step).  The CTA's rows of w stay in shared memory from the first pass to the
// normalisation.  Saves three launches and their fill/drain per step and two passes over w.
// Optional timeline of the fused step (-DLGPU_TRACE, scripts/cgs_trace.py): thread 0 of every CTA stores %globaltimer
// at phase boundaries.  Never compiled into the product library.
#ifdef LGPU_TRACE
constexpr int CTRACE_SLOTS = 16, CTRACE_CTAS = 160;
__device__ unsigned long long g_cgs_trace[CTRACE_CTAS * CTRACE_SLOTS];
__device__ __forceinline__ void cgs_trace_mark(int slot) {
  if (threadIdx.x == 0 && blockIdx.x < CTRACE_CTAS) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_cgs_trace[blockIdx.x * CTRACE_SLOTS + slot] = t;
  }
  __syncwarp();
}
#define CGS_MARK(s) cgs_trace_mark(s)
#else
#define CGS_MARK(s)
#endif

struct CgsArgs {
  BasisLayout L;
  cd* V;
  int ncols;
  int nstages;
  int tiles_max;                 // rows of wkeep / 64
  cd* w;                         // in: OP v_j ; out: the orthogonalised residual
  cd* partial;                   // 3 x gridDim x PSTRIDE
  cd* Hcol;                      // H(0:ncols, j) = h + s
  double* scal;                  // scal[0] = || w ||
  unsigned long long* gbar;      // device-wide barrier counter (monotonic)
  unsigned long long bar_base;   // its value when this launch starts
  int newcol;                    // >= 0: V(:, newcol) = vplain = w / ||w|| and *hsub = ||w||
  cd* vplain;
  cd* hsub;
  // per-chunk completion flags of the kernel that produces w (KrylovWork::wflags), or null
  const unsigned long long* wflags;
  unsigned long long wepoch;
  int wtile_shift, wnchunks;
  int early_trigger;             // LGPU_CGS2_EARLY=1: let the next solve's first kernel start (prologue only) during this step
};

// Device-wide barrier of the consumer threads (pattern of cooperative groups' grid sync): the CTA
// barrier orders every consumer's earlier stores before thread 0's fence + arrival, thread 0's
// acquire load orders the other CTAs' stores before the second CTA barrier.  Data published
// across the barrier is read with ld.global.cg (L2), never through L1.
__device__ __forceinline__ void cgs_grid_barrier(unsigned long long* ctr, unsigned long long target) {
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1ULL);
    unsigned long long v;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
    } while (v < target);
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
}

__device__ __forceinline__ cd ldcg_cd(const cd* p) {
  const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
  return cd{v.x, v.y};
}

// every CTA: hs[c] = sum over CTAs of partial[b][c] in one fixed order.  Thread (c, g) of the 64 x 4
// consumer threads adds the partials of CTAs g, g + 4, ... : all its loads are independent and in
// flight together (one L2 round trip instead of one per partial), consecutive threads read
// consecutive columns of one CTA's row; the four group sums are combined through shared memory.
constexpr int CGS_MAX_GRID = 160;   // CTAs of the fused step kernel (one per SM)
__device__ __forceinline__ void cgs_sum_partials(const cd* partial, int ncols, cd* hs, cd* scratch, int tid) {
  const int c = tid & 63, g = tid >> 6;
  const int nb = static_cast<int>(gridDim.x);
  constexpr int PER = CGS_MAX_GRID / PASS_GROUPS;
  cd x[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int b = g + PASS_GROUPS * k;
    x[k] = (c < ncols && b < nb) ? ldcg_cd(partial + static_cast<size_t>(b) * PSTRIDE + c) : cd{0.0, 0.0};
  }
  cd s0{0.0, 0.0}, s1{0.0, 0.0};
#pragma unroll
  for (int k = 0; k < PER; k += 2) { s0 += x[k]; s1 += x[k + 1]; }
  scratch[g * 64 + c] = s0 + s1;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (tid < ncols) hs[tid] = (scratch[tid] + scratch[64 + tid]) + (scratch[128 + tid] + scratch[192 + tid]);
  asm volatile("bar.sync 1, 256;" ::: "memory");
}

// CPG > 0: the basis has exactly 4 * CPG columns in play as far as the loops are concerned (columns
// ncols .. 4 * CPG - 1 are stale but finite data of the basis, their coefficients are zero and
// their projections are never published): no per-column predicates in the tile loops.
// CPG == 0: any ncols <= KRYLOV_PASS_MAXCOL, predicated.
// How many times slot i % S has been filled before local tile i is streamed in pass `pass` of the
// fused step (passes run up, down, up over the nt tiles of a CTA; the S tiles a pass ends with stay
// in their slots for the next pass).  Its parity is the phase of the slot's "full" barrier.
// x / S for 0 <= x < 8192 and 2 <= S <= 8 without an integer division: M = 65536 / S + 1
struct SmallDiv {
  int S, M;
  __host__ __device__ int div(int x) const { return static_cast<int>((static_cast<unsigned>(x) * static_cast<unsigned>(M)) >> 16); }
  __host__ __device__ int mod(int x) const { return x - div(x) * S; }
};
__device__ __forceinline__ int cgs_fill_number(int pass, int i, int nt, const SmallDiv& d) {
  const int S = d.S, q = d.div(i), sg = i - q * S;
  const int n = d.div(nt - sg + S - 1);       // tiles of this CTA that map to the slot
  if (pass == 1) return q;
  if (pass == 2) return n + d.div(sg + (n - 1) * S - S - i);   // below the slot's resident tile, downwards
  return 2 * n - 1 + q - 1;                   // pass 3: above the slot's resident tile, upwards
}

template <int CPG>
__global__ void __launch_bounds__(PASS_THREADS, 1) krylov_cgs2_kernel(const __grid_constant__ CgsArgs a) {
  constexpr bool EXACT = CPG > 0;
  constexpr int NJ = EXACT ? CPG : PASS_CPG;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const BasisLayout& L = a.L;
  const int ncols = a.ncols, nstages = a.nstages;
  const int ncopy = EXACT ? PASS_GROUPS * CPG : ncols;   // columns a stage holds
  const int sstride = (ncopy + 1) * PASS_T;
  cd* buf = reinterpret_cast<cd*>(smem_raw);                            // [nstages][sstride]
  cd* wkeep = buf + static_cast<size_t>(nstages) * sstride;              // [tiles_max][64]
  cd* part = wkeep + static_cast<size_t>(a.tiles_max) * PASS_T;          // [2][2][4][64]
  cd* hs = part + 4 * PASS_GROUPS * PASS_T;                             // [KRYLOV_PASS_MAXCOL]
  uint64_t* full = reinterpret_cast<uint64_t*>(hs + KRYLOV_PASS_MAXCOL);
  uint64_t* empty = full + nstages;
  __shared__ double red[8];
  __shared__ double s_norm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t0 = static_cast<int>(static_cast<long long>(blockIdx.x) * L.ntiles / gridDim.x);
  const int t1 = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * L.ntiles / gridDim.x);
  const int nt = t1 - t0;
  CGS_MARK(0);
  if (tid == 0) {
    for (int i = 0; i < nstages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  // completion flag of the first tile's chunk: polled by the producer lane while thread 0 sets up the barriers
  int have_chunk = -1;
  if (a.wflags != nullptr && tid == 8 * 32 && nt > 0) {
    have_chunk = min(t0 >> a.wtile_shift, a.wnchunks - 1);
    unsigned long long v;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(a.wflags + have_chunk) : "memory");
    } while (v < a.wepoch);
  }
  __syncthreads();
  if (a.early_trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 8) {
    // ---- producer (one lane).  Local tile i always lives in ring slot i % S, and the passes run
    // over the CTA's tiles in alternating directions (up, down, up): the last S tiles of a pass are
    // the first S of the next one and are still in their slots, so they are neither released nor
    // streamed again (S / nt of the traffic of passes 2 and 3).  Fill number F of a slot waits for
    // release F - 1 of that slot; only the first pass needs w from global memory.
    if (lane == 0) {
      const int S = nstages;
      const SmallDiv sd{S, 65536 / S + 1};
      const uint64_t keep_policy = l2_policy_evict_last();
      const uint32_t vbytes = static_cast<uint32_t>(sizeof(cd) * ncopy * PASS_T);
      auto fill = [&](int i, int F, bool with_w) {
        const int sg = sd.mod(i), t = t0 + i;
        if (F >= 1) mbar_wait(&empty[sg], static_cast<uint32_t>((F - 1) & 1));
        cd* sb = buf + static_cast<size_t>(sg) * sstride;
        const uint32_t wbytes = with_w ? static_cast<uint32_t>(sizeof(cd) * min(PASS_T, L.n - t * PASS_T)) : 0u;
        mbar_expect_tx(&full[sg], vbytes + wbytes);
        // the basis (<= ~100 MB) is re-read by every step and fits the 126 MB L2 next to the upper-level
        // factor records; the stage-0 records of the solve in between stream through with evict-first
        bulk_g2s_hint(sb, a.V + static_cast<size_t>(t) * L.ncv * PASS_T, vbytes, &full[sg], keep_policy);
        if (wbytes) bulk_g2s(sb + ncopy * PASS_T, a.w + static_cast<size_t>(t) * PASS_T, wbytes, &full[sg]);
      };
      // Programmatic dependent launch (krylov_cgs2_step): this CTA may be running while the kernel that produces w
      // is still at work.  Without completion flags nothing is read before that kernel has completed.  With them
      // (the solve's last kernel publishes, chunk by chunk, that its part of w is final) a tile is requested as soon
      // as its chunk is done - the first pass starts while that kernel's last CTAs are still running, and there is
      // no grid-completion latency between the two.  The first flag this lane sees also orders everything it reads
      // after the completion of the whole chain before that chunk (solve kernels, the previous step's launch that
      // wrote the newest basis column): each of them waited for its predecessor's end.
      if (a.wflags == nullptr) asm volatile("griddepcontrol.wait;" ::: "memory");
      else asm volatile("fence.proxy.async;" ::: "memory");
      for (int i = 0; i < nt; ++i) {
        if (a.wflags != nullptr) {
          const int c = min((t0 + i) >> a.wtile_shift, a.wnchunks - 1);
          if (c != have_chunk) {
            unsigned long long v;
            do {
              asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(a.wflags + c) : "memory");
            } while (v < a.wepoch);
            asm volatile("fence.proxy.async;" ::: "memory");   // the bulk copies below read what generic stores wrote
            have_chunk = c;
          }
        }
        fill(i, cgs_fill_number(1, i, nt, sd), true);
      }
      for (int i = nt - S - 1; i >= 0; --i) fill(i, cgs_fill_number(2, i, nt, sd), false);
      for (int i = S; i < nt; ++i) fill(i, cgs_fill_number(3, i, nt, sd), false);
    }
    return;
  }
  // ---- consumers: 64 rows x 4 column groups
  // (griddepcontrol.launch_dependents only behind LGPU_CGS2_EARLY, see the top of the kernel: letting the next solve's
  // first-stage kernel start while this step is still running is worth 0.15 ms per headline step.  It looked
  // non-reproducible with three units in flight until the ring releases got their cross-proxy fence
  // (common.cuh: mbar_release_slot); with the fence 0 of 250 in-flight repetitions differ, profiles/tuning_log_r2.md.)
  // nothing below may precede the kernel before this one - unless w comes with completion flags: then the consumers
  // touch global memory only after the first device-wide barrier, i.e. after every tile of w has been loaded
  if (a.wflags == nullptr) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int cpg = EXACT ? CPG : (ncols + PASS_GROUPS - 1) / PASS_GROUPS;
  const int r = tid & (PASS_T - 1), q = tid >> 6;
  if (EXACT) {   // coefficients of the padding columns stay zero for the whole step
    if (tid < KRYLOV_PASS_MAXCOL) hs[tid] = cd{0.0, 0.0};
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }
  int flip = 0;
  cd acc[NJ];
  cd v[NJ];
  cd* xw = part;   // [8 warps][PASS_CPG] scratch of the row reduction

  const SmallDiv csd{nstages, 65536 / nstages + 1};
  // local tile i of pass `pass` -> registers (see the producer for the slot / residency rules)
  auto load_tile_into = [&](cd (&v)[NJ], int pass, int i, bool want_w, cd& wi) {
    const int S = nstages, sg = csd.mod(i);
    const bool valid = (t0 + i) * PASS_T + r < L.n;
    const bool resident = (pass == 2 && i >= nt - S) || (pass == 3 && i < S);
    if (!resident) mbar_wait(&full[sg], static_cast<uint32_t>(cgs_fill_number(pass, i, nt, csd) & 1));
    const cd* sb = buf + static_cast<size_t>(sg) * sstride;
    if (want_w) wi = valid ? sb[ncopy * PASS_T + r] : cd{0.0, 0.0};
    if (EXACT) {   // rows past the end of the basis are zero in memory
      const cd* col = sb + q * (CPG * PASS_T) + r;
#pragma unroll
      for (int j = 0; j < NJ; ++j) v[j] = col[j * PASS_T];
    } else {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int c = q * cpg + j;
        v[j] = (j < cpg && c < ncols && valid) ? sb[c * PASS_T + r] : cd{0.0, 0.0};
      }
    }
    // the tile now lives in registers: release the slot unless the next pass starts with it
    const bool keep = (pass == 1 && i >= nt - S) || (pass == 2 && i < S);
    if (!keep) mbar_release_slot(&empty[sg], lane);
    else __syncwarp();
  };
  auto load_tile = [&](int pass, int i, bool want_w, cd& wi) { load_tile_into(v, pass, i, want_w, wi); };
  // Two tiles per CTA barrier (passes 2 and 3 of the predicate-free variants): the chain of a tile - registers,
  // partial sums, exchange through shared memory, barrier, corrected w - is latency, not work (2.6 k cycles per tile
  // in pass 2 against 1.2 k cycles per tile of arrival); two independent chains share one barrier.  Same operations
  // per entry in the same order as one tile at a time.
  cd v2[NJ];
  // The four partial sums of row r come from warps (r / 32) + 2 q: rows 0 - 31 and rows 32 - 63 never exchange
  // anything, so each half of the tile has a barrier of its own (128 threads); every scheduler holds one warp of each
  // half, and a half that waits for its exchange leaves the issue slots to the other one.
  auto half_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(2 + (warp & 1)) : "memory"); };
  auto correct_pair = [&](int ia, int ib, cd& wa, cd& wb) {
    cd pa0{0.0, 0.0}, pa1{0.0, 0.0}, pb0{0.0, 0.0}, pb1{0.0, 0.0};
    const cd* hq = hs + q * NJ;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const cd hj = hq[j];
      if (j & 1) { cfma(pa1, v[j], hj); cfma(pb1, v2[j], hj); }
      else { cfma(pa0, v[j], hj); cfma(pb0, v2[j], hj); }
    }
    cd* pp = part + flip * 2 * PASS_GROUPS * PASS_T;
    flip ^= 1;
    pp[q * PASS_T + r] = pa0 + pa1;
    pp[(PASS_GROUPS + q) * PASS_T + r] = pb0 + pb1;
    const cd wolda = wkeep[ia * PASS_T + r], woldb = wkeep[ib * PASS_T + r];
    half_sync();
    const cd* pb = pp + PASS_GROUPS * PASS_T;
    wa = wolda - ((pp[r] + pp[PASS_T + r]) + (pp[2 * PASS_T + r] + pp[3 * PASS_T + r]));
    wb = woldb - ((pb[r] + pb[PASS_T + r]) + (pb[2 * PASS_T + r] + pb[3 * PASS_T + r]));
  };
  // w_r -= sum_c V(r, c) hs[c] from the registers of the four column groups of row r
  auto correct = [&](int i) -> cd {
    cd p0{0.0, 0.0}, p1{0.0, 0.0};
    if (EXACT) {
      const cd* hq = hs + q * CPG;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (j & 1) cfma(p1, v[j], hq[j]); else cfma(p0, v[j], hq[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < NJ; j += 2) {
        if (j < cpg) {
          const int c = q * cpg + j;
          cfma(p0, v[j], hs[min(c, ncols - 1)]);
          cfma(p1, v[j + 1], hs[min(c + 1, ncols - 1)]);
        }
      }
    }
    cd* pp = part + flip * 2 * PASS_GROUPS * PASS_T;
    flip ^= 1;
    pp[q * PASS_T + r] = p0 + p1;
    const cd wold = wkeep[i * PASS_T + r];
    half_sync();
    return wold - ((pp[r] + pp[PASS_T + r]) + (pp[2 * PASS_T + r] + pp[3 * PASS_T + r]));
  };
  auto publish_dots = [&](cd* dst) {   // rows -> one value per (CTA, column)
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      if (EXACT || j < cpg) {
        acc[j].x = warp_sum(acc[j].x);
        acc[j].y = warp_sum(acc[j].y);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) xw[warp * PASS_CPG + j] = acc[j];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    for (int e = tid; e < PASS_GROUPS * cpg; e += 256) {
      const int qq = e / cpg, j = e - qq * cpg;
      const int c = qq * cpg + j;
      if (c < ncols)
        dst[static_cast<size_t>(blockIdx.x) * PSTRIDE + c] = xw[(2 * qq) * PASS_CPG + j] + xw[(2 * qq + 1) * PASS_CPG + j];
    }
  };

  cd* partial1 = a.partial;
  cd* partial2 = a.partial + static_cast<size_t>(gridDim.x) * PSTRIDE;

  // ---- pass 1: h = V^H w
#pragma unroll
  for (int j = 0; j < NJ; ++j) acc[j] = cd{0.0, 0.0};
  for (int i = 0; i < nt; ++i) {
    cd wi{0.0, 0.0};
    load_tile(1, i, true, wi);
#ifdef LGPU_TRACE
    if (i == 0) CGS_MARK(1);
#endif
    if (q == 0) wkeep[i * PASS_T + r] = wi;
#pragma unroll
    for (int j = 0; j < NJ; ++j)
      if (EXACT || j < cpg) cfmac(acc[j], v[j], wi);
  }
  CGS_MARK(2);
  publish_dots(partial1);
  CGS_MARK(3);
  cgs_grid_barrier(a.gbar, a.bar_base + gridDim.x);
  CGS_MARK(4);
  cgs_sum_partials(partial1, ncols, hs, part, tid);
  CGS_MARK(5);
  if (blockIdx.x == 0) for (int c = tid; c < ncols; c += 256) a.Hcol[c] = hs[c];

  // ---- pass 2: w -= V h ; s = V^H w ; || w ||^2
  double nrm = 0.0;
#pragma unroll
  for (int j = 0; j < NJ; ++j) acc[j] = cd{0.0, 0.0};
  int i2 = nt - 1;
  if (EXACT) {
    for (; i2 >= 1; i2 -= 2) {
      cd wa{0.0, 0.0}, wb{0.0, 0.0};
      load_tile_into(v, 2, i2, false, wa);
      load_tile_into(v2, 2, i2 - 1, false, wb);
      correct_pair(i2, i2 - 1, wa, wb);
      if ((t0 + i2) * PASS_T + r >= L.n) wa = cd{0.0, 0.0};
      if (q == 0) {
        wkeep[i2 * PASS_T + r] = wa;
        nrm += abs2(wa);
        wkeep[(i2 - 1) * PASS_T + r] = wb;
        nrm += abs2(wb);
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) cfmac(acc[j], v[j], wa);
#pragma unroll
      for (int j = 0; j < NJ; ++j) cfmac(acc[j], v2[j], wb);
    }
  }
  for (int i = i2; i >= 0; --i) {
    cd wi{0.0, 0.0};
    load_tile(2, i, false, wi);
    wi = correct(i);
    if ((t0 + i) * PASS_T + r >= L.n) wi = cd{0.0, 0.0};
    if (q == 0) { wkeep[i * PASS_T + r] = wi; nrm += abs2(wi); }
#pragma unroll
    for (int j = 0; j < NJ; ++j)
      if (EXACT || j < cpg) cfmac(acc[j], v[j], wi);
  }
  CGS_MARK(6);
  publish_dots(partial2);
  nrm = warp_sum(nrm);
  if (lane == 0) red[warp] = nrm;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (tid == 0) partial2[static_cast<size_t>(blockIdx.x) * PSTRIDE + KRYLOV_MAXCOL] = cd{red[0] + red[1], 0.0};
  CGS_MARK(7);
  cgs_grid_barrier(a.gbar, a.bar_base + 2ull * gridDim.x);
  CGS_MARK(8);
  // the per-CTA squared norms travel with the partials (the loads are issued first and are in flight during the
  // summation of the partials: one L2 round trip for both)
  double x[CGS_MAX_GRID / 32];
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < CGS_MAX_GRID / 32; ++k) {
      const unsigned int b = lane + 32 * k;
      x[k] = b < gridDim.x ? ldcg_cd(partial2 + static_cast<size_t>(b) * PSTRIDE + KRYLOV_MAXCOL).x : 0.0;
    }
  }
  cgs_sum_partials(partial2, ncols, hs, part, tid);
  if (blockIdx.x == 0) for (int c = tid; c < ncols; c += 256) a.Hcol[c] = a.Hcol[c] + hs[c];
  // The norm of the final residual without a third device-wide round: the basis is orthonormal,
  // so || w - V s ||^2 = || w ||^2 - || s ||^2, and after the first correction || s || is at
  // rounding level of || w || (no cancellation).  Every CTA evaluates it in the same order.
  if (warp == 0) {
    double wn2 = 0.0;
#pragma unroll
    for (int k = 0; k < CGS_MAX_GRID / 32; ++k) wn2 += x[k];
    wn2 = warp_sum(wn2);
    double ss = 0.0;
    for (int c = lane; c < ncols; c += 32) ss += abs2(hs[c]);
    ss = warp_sum(ss);
    if (lane == 0) s_norm = sqrt(fmax(wn2 - ss, 0.0));
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  const double rnorm = s_norm;
  if (blockIdx.x == 0 && tid == 0) {
    a.scal[0] = rnorm;
    if (a.hsub) *a.hsub = cd{rnorm, 0.0};   // H(j + 1, j); for the last column of a full basis: the slot behind the matrix
  }

  CGS_MARK(9);
  // ---- pass 3: w -= V s, and the next basis vector V(:, newcol) = vplain = w / ||w|| on the way
  const double inv = 1.0 / rnorm;
  auto emit = [&](int i, cd wi) {
    if (q == 0) {
      const int gi = (t0 + i) * PASS_T + r;
      if (gi < L.n) {
        a.w[gi] = wi;
        if (a.newcol >= 0) {
          const cd x = wi * inv;
          a.V[(static_cast<size_t>(t0 + i) * L.ncv + a.newcol) * PASS_T + r] = x;
          a.vplain[gi] = x;
        }
      }
    }
  };
  int i3 = 0;
  if (EXACT) {
    for (; i3 + 1 < nt; i3 += 2) {
      cd wa{0.0, 0.0}, wb{0.0, 0.0};
      load_tile_into(v, 3, i3, false, wa);
      load_tile_into(v2, 3, i3 + 1, false, wb);
      correct_pair(i3, i3 + 1, wa, wb);
      emit(i3, wa);
      emit(i3 + 1, wb);
    }
  }
  for (int i = i3; i < nt; ++i) {
    cd wi{0.0, 0.0};
    load_tile(3, i, false, wi);
    wi = correct(i);
    emit(i, wi);
  }
  CGS_MARK(10);
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

// Restart / extraction GEMM, one thread per row: the row of V is pulled into registers with NK
// independent coalesced loads, every output column is then a dot product against a column of Q
// broadcast from shared memory (one wavefront per FMA group), so the kernel is bounded by the
// FP64 pipe and HBM instead of shared-memory traffic; safe in place for any nc (the row is in
// registers before the first store).  NK = 40 covers ARPACK's default ncv = 2 nev of this path.
template <int NK>
__global__ void __launch_bounds__(128)
basis_gemm_rows_kernel(BasisLayout L, const cd* V, int nk, const cd* __restrict__ Q, int ldq, int nc,
                       cd* Out, int out_plain_ld) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* Qs = reinterpret_cast<cd*>(smem_raw);   // [nc][nk]
  for (int e = threadIdx.x; e < nk * nc; e += 128) {
    const int j = e % nk, c = e / nk;
    Qs[e] = Q[static_cast<size_t>(c) * ldq + j];
  }
  const int i = blockIdx.x * 128 + threadIdx.x;
  const int ic = min(i, L.n - 1);
  const int t = ic / L.T, r = ic - t * L.T;
  const cd* row = V + static_cast<size_t>(t) * L.ncv * L.T + r;   // column j at row[j * T]
  cd v[NK];
#pragma unroll
  for (int j = 0; j < NK; ++j) v[j] = j < nk ? row[static_cast<size_t>(j) * L.T] : cd{0.0, 0.0};
  __syncthreads();
  if (i >= L.n) return;
  for (int c = 0; c < nc; ++c) {
    const cd* q = Qs + c * nk;
    cd a0{0.0, 0.0}, a1{0.0, 0.0}, a2{0.0, 0.0}, a3{0.0, 0.0};
#pragma unroll
    for (int j = 0; j < NK; j += 4) {
      if (j < nk) cfma(a0, v[j], q[j]);
      if (j + 1 < nk) cfma(a1, v[j + 1], q[j + 1]);
      if (j + 2 < nk) cfma(a2, v[j + 2], q[j + 2]);
      if (j + 3 < nk) cfma(a3, v[j + 3], q[j + 3]);
    }
    const cd acc = (a0 + a1) + (a2 + a3);
    if (out_plain_ld > 0) Out[static_cast<size_t>(c) * out_plain_ld + i] = acc;
    else Out[(static_cast<size_t>(t) * L.ncv + c) * L.T + r] = acc;
  }
}

// Restart / extraction GEMM on the FP64 tensor cores (mma.sync.m8n8k4.f64): Out(64 rows of a tile, nc) = V(tile, nk) Q.
// The DMMA pipe is no faster than DFMA on this GPU (37 vs 35 TFLOP/s, profiles/fp64_pipe_r1.md) - the point is the
// instruction count: the row-per-thread kernel above issues one broadcast shared-memory load per complex FMA and
// reaches ~36 % of the FP64 peak (99 us for 1.2 GFLOP); here a warp owns 8 rows, its A fragments (one complex value
// per lane and k-step) are reused for every block of 8 output columns, and one shared-memory load feeds four MMAs
// (real / imaginary x real / imaginary).  Persistent CTAs, Q staged once per CTA, one 64-row tile (contiguous in the
// basis layout) at a time; in place is safe (the tile is in shared memory before the first store).
//   smem: A [nkp][A_LD] complex (A_LD = 66: the 8 lanes of a quarter warp hit 8 different 16-byte bank groups),
//         Q [ncp][nkp + 4] complex (same reason), nkp / ncp = nk / nc rounded up to 4 / 8
constexpr int GEMM_A_LD = PASS_T + 2;
__device__ __forceinline__ void dmma_884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NBLK>   // blocks of 8 output columns
__global__ void __launch_bounds__(256)
basis_gemm_mma_kernel(BasisLayout L, const cd* V, int nk, const cd* __restrict__ Q, int ldq, int nc, cd* Out,
                      int out_plain_ld) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nkp = (nk + 3) & ~3, qld = nkp + 4;
  cd* As = reinterpret_cast<cd*>(smem_raw);          // [nkp][GEMM_A_LD]
  cd* Qs = As + nkp * GEMM_A_LD;                     // [8 NBLK][qld]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < 8 * NBLK * qld; e += 256) {
    const int c = e / qld, j = e - c * qld;
    Qs[e] = (c < nc && j < nk) ? Q[static_cast<size_t>(c) * ldq + j] : cd{0.0, 0.0};
  }
  for (int e = tid; e < (nkp - nk) * PASS_T; e += 256) As[(nk + e / PASS_T) * GEMM_A_LD + (e % PASS_T)] = cd{0.0, 0.0};
  const int g = lane >> 2, q4 = lane & 3;            // fragment coordinates: row / column group, k index
  for (int t = blockIdx.x; t < L.ntiles; t += gridDim.x) {
    const cd* src = V + static_cast<size_t>(t) * L.ncv * PASS_T;
    __syncthreads();                                 // the previous tile's fragments have been read
    for (int e = tid; e < nk * PASS_T; e += 256) {
      const double2 v = *reinterpret_cast<const double2*>(src + e);
      As[(e >> 6) * GEMM_A_LD + (e & 63)] = cd{v.x, v.y};
    }
    __syncthreads();
    double cr[NBLK][2], ci[NBLK][2];
#pragma unroll
    for (int b = 0; b < NBLK; ++b) { cr[b][0] = cr[b][1] = ci[b][0] = ci[b][1] = 0.0; }
    const cd* arow = As + q4 * GEMM_A_LD + 8 * warp + g;
    const cd* qrow = Qs + g * qld + q4;
    for (int ks = 0; ks < nkp; ks += 4) {
      const cd a = arow[ks * GEMM_A_LD];
      const double nai = -a.y;
#pragma unroll
      for (int b = 0; b < NBLK; ++b) {
        const cd q = qrow[8 * b * qld + ks];
        dmma_884(cr[b][0], cr[b][1], a.x, q.x);
        dmma_884(cr[b][0], cr[b][1], nai, q.y);
        dmma_884(ci[b][0], ci[b][1], a.x, q.y);
        dmma_884(ci[b][0], ci[b][1], a.y, q.x);
      }
    }
    // C fragment: row g of the warp's 8, columns 8 b + 2 q4 + {0, 1}
    const int r = 8 * warp + g, gi = t * PASS_T + r;
#pragma unroll
    for (int b = 0; b < NBLK; ++b) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = 8 * b + 2 * q4 + h;
        if (c < nc && gi < L.n) {
          const cd v{cr[b][h], ci[b][h]};
          if (out_plain_ld > 0) Out[static_cast<size_t>(c) * out_plain_ld + gi] = v;
          else Out[(static_cast<size_t>(t) * L.ncv + c) * PASS_T + r] = v;
        }
      }
    }
  }
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

// r = a*r + b*V(:, col) and scal[0] = ||r|| in one launch (the residual of an implicit restart):
// CTA partials of the squared norm, the last CTA to finish sums them in a fixed order
__global__ void __launch_bounds__(256)
vec_axpby_basis_norm_kernel(BasisLayout L, cd a, cd* __restrict__ r, cd b, const cd* __restrict__ V, int col,
                            cd* __restrict__ partial, double* scal, unsigned int* ticket) {
  __shared__ double red[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double nn = 0.0;
  for (int i = blockIdx.x * 256 + tid; i < L.n; i += gridDim.x * 256) {
    const cd v = a * r[i] + b * V[basis_off(L, i, col)];
    r[i] = v;
    nn += abs2(v);
  }
  nn = warp_sum(nn);
  if (lane == 0) red[warp] = nn;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[static_cast<size_t>(blockIdx.x) * PSTRIDE] = cd{t, 0.0};
  }
  if (last_block_done(ticket, tid == 0)) {
    if (warp == 0) {
      double t = 0.0;
      for (unsigned int bb = lane; bb < gridDim.x; bb += 32) t += partial[static_cast<size_t>(bb) * PSTRIDE].x;
      t = warp_sum(t);
      if (lane == 0) scal[0] = sqrt(t);
    }
    __syncthreads();
    if (tid == 0) *ticket = 0u;
  }
}

// out[0] = x^H s, out[1] = x^H r (zdotc), out[2] = (||x||^2, 0): CTA partials, the last CTA sums
// them in a fixed order
__global__ void __launch_bounds__(256)
vec_dot2_kernel(int n, const cd* __restrict__ x, const cd* __restrict__ sv, const cd* __restrict__ rv,
                cd* __restrict__ partial, cd* out, unsigned int* ticket) {
  __shared__ cd red[3][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  cd a0{0.0, 0.0}, a1{0.0, 0.0};
  double nn = 0.0;
  for (int i = blockIdx.x * 256 + tid; i < n; i += gridDim.x * 256) {
    const cd xi = x[i];
    cfmac(a0, xi, sv[i]);
    cfmac(a1, xi, rv[i]);
    nn += abs2(xi);
  }
  a0.x = warp_sum(a0.x); a0.y = warp_sum(a0.y);
  a1.x = warp_sum(a1.x); a1.y = warp_sum(a1.y);
  nn = warp_sum(nn);
  if (lane == 0) { red[0][warp] = a0; red[1][warp] = a1; red[2][warp] = cd{nn, 0.0}; }
  __syncthreads();
  if (tid < 3) {
    cd t{0.0, 0.0};
    for (int w = 0; w < 8; ++w) t += red[tid][w];
    partial[static_cast<size_t>(blockIdx.x) * PSTRIDE + tid] = t;
  }
  if (last_block_done(ticket, tid < 3)) {
    if (warp < 3) {
      cd t{0.0, 0.0};
      for (unsigned int b = lane; b < gridDim.x; b += 32) t += partial[static_cast<size_t>(b) * PSTRIDE + warp];
      t.x = warp_sum(t.x);
      t.y = warp_sum(t.y);
      if (lane == 0) out[warp] = t;
    }
    __syncthreads();
    if (tid == 0) *ticket = 0u;
  }
}

__global__ void __launch_bounds__(256)
vec_axpby_kernel(int n, cd a, cd* __restrict__ r, cd b, const cd* __restrict__ v) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) r[i] = a * r[i] + b * v[i];
}

}  // namespace

#ifdef LGPU_TRACE
extern "C" int lgpu_debug_cgs_trace(unsigned long long* out) {
  return static_cast<int>(cudaMemcpyFromSymbol(out, g_cgs_trace, sizeof(unsigned long long) * CTRACE_CTAS * CTRACE_SLOTS));
}
#endif

BasisLayout make_basis_layout(int n, int ncv) {
  BasisLayout L{};
  L.n = n;
  L.ncv = ncv;
  L.T = KRYLOV_TILE;
  L.ntiles = (n + L.T - 1) / L.T;
  return L;
}

static int sm_count() { return device_sm_count(); }
static size_t pass_smem(int ncols, int nstages) {
  return sizeof(cd) * (static_cast<size_t>(nstages) * (ncols + 1) * PASS_T + 2 * PASS_GROUPS * PASS_T +
                       KRYLOV_PASS_MAXCOL) + 16 * nstages;
}
// columns [c0, c0 + ncols) of the basis in one launch (ncols <= KRYLOV_PASS_MAXCOL)
template <bool UPDATE, bool DOTS>
static void launch_pass(const BasisLayout& L, const cd* V, int c0, int ncols, cd* w, const KrylovWork& work,
                        cd* Hcol, int accumulate, cudaStream_t stream) {
  static PerDeviceOnce once;
  once.run([] {
    CUDA_CHECK(cudaFuncSetAttribute(krylov_pass_kernel<UPDATE, DOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    220 * 1024));
  });
  int nstages = 6;
  while (nstages > 2 && pass_smem(ncols, nstages) > 200 * 1024) --nstages;
  const int grid = std::max(1, std::min(sm_count(), L.ntiles));
  krylov_pass_kernel<UPDATE, DOTS><<<grid, PASS_THREADS, pass_smem(ncols, nstages), stream>>>(
      L, V + static_cast<size_t>(c0) * PASS_T, ncols, nstages, w, work.hwork + c0, work.partial, work.hwork + c0,
      Hcol ? Hcol + c0 : nullptr, accumulate, work.scal, work.ticket);
}

void krylov_dots(const BasisLayout& L, const cd* V, int ncols, const cd* w, const KrylovWork& work,
                 cd* Hcol, int accumulate, cudaStream_t stream, LaunchLog* log) {
  log->begin(LK_DOTS, 16.0 * L.n * (ncols + 1));
  for (int c0 = 0; c0 < ncols; c0 += KRYLOV_PASS_MAXCOL) {
    launch_pass<false, true>(L, V, c0, std::min(KRYLOV_PASS_MAXCOL, ncols - c0), const_cast<cd*>(w), work, Hcol,
                             accumulate, stream);
    log->launches += 1;
  }
  log->end();
  CUDA_CHECK(cudaGetLastError());
}

void krylov_update_dots(const BasisLayout& L, const cd* V, int ncols, cd* w, const KrylovWork& work,
                        cd* Hcol, int accumulate, cudaStream_t stream, LaunchLog* log) {
  if (ncols > KRYLOV_PASS_MAXCOL) {   // wide basis: the projection needs the fully corrected w
    krylov_update(L, V, ncols, w, work, stream, log);
    krylov_dots(L, V, ncols, w, work, Hcol, accumulate, stream, log);
    return;
  }
  log->begin(LK_DOTS, 16.0 * L.n * (ncols + 2));
  launch_pass<true, true>(L, V, 0, ncols, w, work, Hcol, accumulate, stream);
  log->end();
  log->launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

void krylov_update(const BasisLayout& L, const cd* V, int ncols, cd* w, const KrylovWork& work,
                   cudaStream_t stream, LaunchLog* log) {
  log->begin(LK_UPDATE, 16.0 * L.n * (ncols + 2));
  int c0 = 0;
  do {   // the last launch leaves ||w|| in scal[0]
    launch_pass<true, false>(L, V, c0, std::min(KRYLOV_PASS_MAXCOL, ncols - c0), w, work, nullptr, 0, stream);
    log->launches += 1;
    c0 += KRYLOV_PASS_MAXCOL;
  } while (c0 < ncols);
  log->end();
  CUDA_CHECK(cudaGetLastError());
}

// dynamic shared memory the fused step may use: the 227 KB of an SM less its ~1.1 KB of static data
constexpr size_t CGS2_SMEM_MAX = 225 * 1024;
static size_t cgs2_smem(int ncols, int nstages, int tiles_max) {
  return sizeof(cd) * (static_cast<size_t>(nstages) * (ncols + 1) * PASS_T + static_cast<size_t>(tiles_max) * PASS_T +
                       4 * PASS_GROUPS * PASS_T + KRYLOV_PASS_MAXCOL) + 16 * nstages;
}

template <int CPG>
static void launch_cgs2(const CgsArgs& a, int grid, size_t smem, cudaStream_t stream, bool pdl) {
  static PerDeviceOnce once;
  once.run([] {
    CUDA_CHECK(cudaFuncSetAttribute(krylov_cgs2_kernel<CPG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(CGS2_SMEM_MAX)));
  });
  // Cooperative (all CTAs co-resident: they meet at device-wide barriers) and, for a context that has the GPU to
  // itself, programmatic: the CTAs start while the last solve kernel is still running, initialise their barriers and
  // stream their first basis tiles, and block in griddepcontrol.wait until w is complete.  A context that shares the
  // GPU (SM cap set) keeps the plain launch: its admission ticket counts the solve kernels OR this grid, not both.
  static const bool pdl_env = [] { const char* e = std::getenv("LGPU_PDL"); return !(e && e[0] == '0'); }();
  static bool pdl_ok = true;   // cleared if the driver rejects the attribute pair
  if (a.wflags != nullptr) {
    // With completion flags the CTAs must be able to start while the solve's last kernel is still running: a
    // cooperative grid is started only once ALL its CTAs fit, i.e. after that kernel has drained (timeline: all 148
    // CTAs start 4 us after its last CTA ends).  A plain programmatic launch instead: the CTAs take the SMs one by
    // one as they become free.  They are all resident in the end - this context is alone on the GPU (its admission
    // ticket covers every SM), there is one CTA per SM, and the kernel they wait for never waits for them.
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(PASS_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, krylov_cgs2_kernel<CPG>, a));
    return;
  }
  if (pdl && pdl_env && pdl_ok) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(PASS_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 2;
    const cudaError_t rc = cudaLaunchKernelEx(&cfg, krylov_cgs2_kernel<CPG>, a);
    if (rc == cudaSuccess) return;
    (void)cudaGetLastError();
    pdl_ok = false;
  }
  void* args[] = {const_cast<CgsArgs*>(&a)};
  CUDA_CHECK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(krylov_cgs2_kernel<CPG>), dim3(grid), dim3(PASS_THREADS),
                                         args, smem, stream));
}

int krylov_cgs2_grid(int ntiles, int grid_cap) {
  const int sms = grid_cap > 0 ? std::min(grid_cap, sm_count()) : sm_count();
  return std::max(1, std::min(sms, ntiles));
}

bool krylov_cgs2_step(const BasisLayout& L, cd* V, int ncols, cd* w, const KrylovWork& work, cd* Hcol,
                      int newcol, cd* vplain, cd* hsub, cudaStream_t stream, LaunchLog* log) {
  static const bool enabled = [] { const char* e = std::getenv("LGPU_CGS2_FUSED"); return !(e && e[0] == '0'); }();
  static const bool exact_ok = [] { const char* e = std::getenv("LGPU_CGS2_EXACT"); return !(e && e[0] == '0'); }();
  const int grid = krylov_cgs2_grid(L.ntiles, work.grid_cap);
  const int tiles_max = (L.ntiles + grid - 1) / grid;
  // columns-per-group variants without per-column predicates: they stream 4 * cpg <= ncv columns
  const int cpg = (ncols + PASS_GROUPS - 1) / PASS_GROUPS;
  const bool exact = exact_ok && cpg >= 5 && cpg <= 10 && PASS_GROUPS * cpg <= L.ncv;
  const int ncopy = exact ? PASS_GROUPS * cpg : ncols;
  int nstages = 5;
  while (nstages > 2 && cgs2_smem(ncopy, nstages, tiles_max) > CGS2_SMEM_MAX) --nstages;
  if (!enabled || ncols < 1 || ncols > KRYLOV_PASS_MAXCOL || work.gbar == nullptr || grid > CGS_MAX_GRID ||
      cgs2_smem(ncopy, nstages, tiles_max) > CGS2_SMEM_MAX)
    return false;
  CgsArgs a{};
  a.L = L; a.V = V; a.ncols = ncols; a.nstages = nstages; a.tiles_max = tiles_max; a.w = w;
  a.partial = work.partial; a.Hcol = Hcol; a.scal = work.scal; a.gbar = work.gbar;
  a.bar_base = *work.gbar_count;
  *work.gbar_count += 2ull * grid;
  a.newcol = newcol; a.vplain = vplain; a.hsub = hsub;
  static const bool early = [] { const char* e = std::getenv("LGPU_CGS2_EARLY"); return e && e[0] == '1'; }();
  a.early_trigger = early ? 1 : 0;
  // completion flags only for a context alone on the GPU, with programmatic launches on, one CTA per SM
  static const bool flags_env = [] {
    const char* e = std::getenv("LGPU_CGS2_FLAGS"); const char* p = std::getenv("LGPU_PDL");
    return !(e && e[0] == '0') && !(p && p[0] == '0');
  }();
  if (flags_env && work.wflags != nullptr && work.grid_cap <= 0 && grid <= sm_count()) {
    a.wflags = work.wflags; a.wepoch = work.wepoch; a.wtile_shift = work.wtile_shift; a.wnchunks = work.wnchunks;
  }
  const size_t smem = cgs2_smem(ncopy, nstages, tiles_max);
  // one orthogonalisation at basis size j = two passes, 2 (j + 2) 256 G bytes
  log->begin(LK_CGS2, 16.0 * L.n * (2.0 * ncols + 4.0));
  const bool pdl = work.grid_cap <= 0;
  switch (exact ? cpg : 0) {
    case 5: launch_cgs2<5>(a, grid, smem, stream, pdl); break;
    case 6: launch_cgs2<6>(a, grid, smem, stream, pdl); break;
    case 7: launch_cgs2<7>(a, grid, smem, stream, pdl); break;
    case 8: launch_cgs2<8>(a, grid, smem, stream, pdl); break;
    case 9: launch_cgs2<9>(a, grid, smem, stream, pdl); break;
    case 10: launch_cgs2<10>(a, grid, smem, stream, pdl); break;
    default: launch_cgs2<0>(a, grid, smem, stream, pdl); break;
  }
  log->end();
  log->launches += 1;
  return true;
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
  static PerDeviceOnce once;
  once.run([] {
    CUDA_CHECK(cudaFuncSetAttribute(basis_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  });
  log->begin(LK_GEMM, 16.0 * L.n * (nk + nc));
  static const bool rows_path = [] { const char* e = std::getenv("LGPU_GEMM_ROWS"); return !(e && e[0] == '0'); }();
  static const bool mma_path = [] { const char* e = std::getenv("LGPU_GEMM_MMA"); return !(e && e[0] == '0'); }();
  const int nkp = (nk + 3) & ~3, nblk = (nc + 7) / 8;
  const size_t mma_smem = sizeof(cd) * (static_cast<size_t>(nkp) * GEMM_A_LD + static_cast<size_t>(8 * nblk) * (nkp + 4));
  if (mma_path && L.T == PASS_T && nblk >= 1 && nblk <= 5 && nk <= 64 && mma_smem <= 100 * 1024) {
    static PerDeviceOnce once_mma;
    once_mma.run([] {
      CUDA_CHECK(cudaFuncSetAttribute(basis_gemm_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      CUDA_CHECK(cudaFuncSetAttribute(basis_gemm_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      CUDA_CHECK(cudaFuncSetAttribute(basis_gemm_mma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      CUDA_CHECK(cudaFuncSetAttribute(basis_gemm_mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      CUDA_CHECK(cudaFuncSetAttribute(basis_gemm_mma_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    });
    const int per_sm = std::max(1, std::min(3, static_cast<int>((220 * 1024) / (mma_smem + 1024))));
    const int grid = std::max(1, std::min(L.ntiles, per_sm * sm_count()));
    switch (nblk) {
      case 1: basis_gemm_mma_kernel<1><<<grid, 256, mma_smem, stream>>>(L, V, nk, Q, ldq, nc, Out, out_plain_ld); break;
      case 2: basis_gemm_mma_kernel<2><<<grid, 256, mma_smem, stream>>>(L, V, nk, Q, ldq, nc, Out, out_plain_ld); break;
      case 3: basis_gemm_mma_kernel<3><<<grid, 256, mma_smem, stream>>>(L, V, nk, Q, ldq, nc, Out, out_plain_ld); break;
      case 4: basis_gemm_mma_kernel<4><<<grid, 256, mma_smem, stream>>>(L, V, nk, Q, ldq, nc, Out, out_plain_ld); break;
      default: basis_gemm_mma_kernel<5><<<grid, 256, mma_smem, stream>>>(L, V, nk, Q, ldq, nc, Out, out_plain_ld); break;
    }
  } else if (rows_path && nk <= 40 && sizeof(cd) * nk * nc <= 48 * 1024)
    basis_gemm_rows_kernel<40><<<(L.n + 127) / 128, 128, sizeof(cd) * nk * nc, stream>>>(L, V, nk, Q, ldq, nc, Out,
                                                                                         out_plain_ld);
  else
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

void vec_axpby_basis_norm(const BasisLayout& L, cd a, cd* r, cd b, const cd* V, int col, const KrylovWork& work,
                          cudaStream_t stream, LaunchLog* log) {
  log->begin(LK_OTHER, 48.0 * L.n);
  const int grid = std::max(1, std::min(2 * sm_count(), (L.n + 255) / 256));
  vec_axpby_basis_norm_kernel<<<grid, 256, 0, stream>>>(L, a, r, b, V, col, work.partial, work.scal, work.ticket);
  log->end();
  log->launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

void vec_dot2(int n, const cd* x, const cd* sv, const cd* rv, const KrylovWork& work, cudaStream_t stream,
              LaunchLog* log) {
  log->begin(LK_OTHER, 48.0 * n);
  const int grid = std::max(1, std::min(sm_count(), (n + 255) / 256));
  vec_dot2_kernel<<<grid, 256, 0, stream>>>(n, x, sv, rv, work.partial, work.hwork, work.ticket);
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
