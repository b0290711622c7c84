// K2 / K2' / K3a — see slu.cuh for the algorithm and the reference call sites replaced.
#include "slu.cuh"

#include <algorithm>
#include <cstdlib>

namespace lgpu {

// ============================================================================= plan
SluPlan make_slu_plan(int n, int first_stage_mu, int next_stage_mu, int top_max_rows) {
  first_stage_mu = std::max(1, std::min(first_stage_mu, 6));
  next_stage_mu = std::max(1, std::min(next_stage_mu, 6));
  top_max_rows = std::max(2, std::min(top_max_rows, 64));
  SluPlan p;
  p.n = n;
  p.n_pad = n + (n & 1);
  p.K = p.n_pad / 2;
  const int m0 = p.K - 1;
  p.top_size = m0 >= 1 ? 64 : 32;
  size_t pairs = 0, rows = 0;
  int m = m0;
  while (m > 1) {
    SluLevel lv{};
    lv.m = m;
    lv.npairs = m / 2;
    lv.off_pairs = pairs; pairs += lv.npairs;
    lv.off_rows = rows; rows += m;
    p.levels.push_back(lv);
    m = (m + 1) / 2;
  }
  p.off_rows_final = rows;
  rows += 1;
  p.pair_records = pairs;
  p.work_rows = rows;
  const int nl = static_cast<int>(p.levels.size());
  int l0 = 0;
  size_t rhs = 0;
  bool first = true;
  while (true) {
    const int ml = l0 < nl ? p.levels[l0].m : (m0 >= 1 ? 1 : 0);
    SluStage st{};
    st.l0 = l0;
    st.m0 = ml;
    st.off_fin = rhs;
    if (ml <= top_max_rows || l0 >= nl) {
      st.mu = nl - l0;
      st.nchunks = 1;
      rhs += ml;
      p.stages.push_back(st);
      break;
    }
    st.mu = std::min(first ? first_stage_mu : next_stage_mu, nl - l0);
    const int C = 1 << st.mu;
    st.nchunks = (ml + C - 1) / C;
    rhs += ml;
    p.stages.push_back(st);
    l0 += st.mu;
    first = false;
  }
  p.rhs_vecs = rhs + 1;
  return p;
}

// ============================================================================ device
namespace {

__device__ __forceinline__ cd ldg_cd(const cd* p) {
  const double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return cd{v.x, v.y};
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

__device__ __forceinline__ int tri_lo_off(int c) { return c * SB - (c * (c - 1)) / 2; }
__device__ __forceinline__ int tri_up_off(int k) { return (k * (k + 1)) / 2; }

// Programmatic dependent launch: wait = the kernels before this one on the stream have completed and
// their writes are visible (no-op for a plain launch); launch_dependents = the next kernel's CTAs may be
// scheduled as soon as every CTA of this grid got here (they run their prologue and block in wait).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Optional timeline of the solve kernels (-DLGPU_TRACE, scripts/solve_trace.py): thread 0 of every CTA stores
// %globaltimer at phase boundaries.  Never compiled into the product library.
#ifdef LGPU_TRACE
constexpr int TRACE_SLOTS = 16, TRACE_CTAS = 512;
__device__ unsigned long long g_trace[3][TRACE_CTAS * TRACE_SLOTS];   // [0] forward stage 0, [1] upper, [2] backward stage 0
__device__ __forceinline__ void trace_mark(int which, int slot) {
  if (threadIdx.x == 0 && blockIdx.x < TRACE_CTAS) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_trace[which][blockIdx.x * TRACE_SLOTS + slot] = t;
  }
  __syncwarp();   // thread 0 must rejoin its warp here: a warp left diverged takes the slow path of every shuffle
}
__device__ __forceinline__ void trace_value(int which, int slot, unsigned long long v) {
  if (threadIdx.x == 0 && blockIdx.x < TRACE_CTAS) g_trace[which][blockIdx.x * TRACE_SLOTS + slot] = v;
}
__device__ __forceinline__ unsigned trace_smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#define TRACE_MARK(w, s) trace_mark(w, s)
#define TRACE_VALUE(w, s, v) trace_value(w, s, v)
#else
#define TRACE_MARK(w, s)
#define TRACE_VALUE(w, s, v)
#endif

constexpr int MAX_STAGE_LEVELS = 14;
constexpr int MAX_FUSED = 6;          // stages one cooperative launch can chain (incl. the top stage)

struct LevelRef {
  int m;
  size_t off_pairs;
};

struct StageArgs {
  int l0, mu, m0, nchunks;
  int n, n_pad, K, top_size;
  LevelRef lv[MAX_STAGE_LEVELS];
  const cd* pairs;
  const cd* top;
  const cd* b;       // original right-hand side (n x 16)
  const cd* fin;     // stage input rows (m0 x 32)
  cd* fout;          // stage output rows (nchunks x 32)
  cd* gvec;
  cd* xv;            // solution, K x 32
  // mailbox hand-off inside the fused launch (null / 0 elsewhere)
  int poll_in;       // fin is a mailbox: poll every entry until it is no longer the all-ones NaN
  cd* fout_clear;    // the other parity's copy of fout: receives the all-ones pattern
  cd* zbox;          // this solve's copy of the solved super nodes (K x 32), written next to xv
  cd* zbox_clear;    // the other parity's copy
  int poll_z;        // boundary unknowns of a chunk come from zbox (polled) instead of xv
  RhsEll ell;        // ell.x != nullptr: the right-hand side is B v, formed on the fly (b is not read)
  int fin_is_rhs;    // first stage: fin points into the right-hand side itself
  unsigned long long* done_flags;   // backward first stage: SolveSignal::flags (or null)
  unsigned long long done_epoch;
};

// entry idx of the right-hand side
__device__ __forceinline__ cd rhs_entry(const StageArgs& a, size_t idx) {
  if (a.ell.x == nullptr) {
    const double2 v = __ldcg(reinterpret_cast<const double2*>(a.b + idx));
    return cd{v.x, v.y};
  }
  double v[8];
  int c[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const bool on = k < a.ell.width;
    v[k] = on ? __ldg(a.ell.val + static_cast<size_t>(k) * a.ell.rows + idx) : 0.0;
    c[k] = on ? __ldg(a.ell.col + static_cast<size_t>(k) * a.ell.rows + idx) : 0;
  }
  cd acc{0.0, 0.0};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < a.ell.width) {
      const double2 t = __ldg(reinterpret_cast<const double2*>(a.ell.x + c[k]));
      acc.x = fma(v[k], t.x, acc.x);
      acc.y = fma(v[k], t.y, acc.y);
    }
  }
  return acc;
}

// two entries at once: all loads of a phase issued before any is used.  (Loading the static values / columns of B
// before griddepcontrol.wait was tried: 48 more live registers under the kernel's 72-register cap spill, first
// stage 33.2 -> 37.4 us.)
__device__ __forceinline__ void rhs_entry2(const StageArgs& a, size_t i0, size_t i1, cd& r0, cd& r1) {
  double v0[8], v1[8];
  int c0[8], c1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const bool on = k < a.ell.width;
    const size_t off = static_cast<size_t>(k) * a.ell.rows;
    v0[k] = on ? __ldg(a.ell.val + off + i0) : 0.0;
    v1[k] = on ? __ldg(a.ell.val + off + i1) : 0.0;
    c0[k] = on ? __ldg(a.ell.col + off + i0) : 0;
    c1[k] = on ? __ldg(a.ell.col + off + i1) : 0;
  }
  double2 t0[8], t1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const bool on = k < a.ell.width;
    t0[k] = on ? __ldg(reinterpret_cast<const double2*>(a.ell.x + c0[k])) : make_double2(0.0, 0.0);
    t1[k] = on ? __ldg(reinterpret_cast<const double2*>(a.ell.x + c1[k])) : make_double2(0.0, 0.0);
  }
  r0 = cd{0.0, 0.0};
  r1 = cd{0.0, 0.0};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < a.ell.width) {
      r0.x = fma(v0[k], t0[k].x, r0.x); r0.y = fma(v0[k], t0[k].y, r0.y);
      r1.x = fma(v1[k], t1[k].x, r1.x); r1.y = fma(v1[k], t1[k].y, r1.y);
    }
  }
}

// ---- solve kernels ---------------------------------------------------------------------------
// A CTA reduces a chunk of 2^mu rows level by level.  The factor records a chunk needs do not
// depend on the right-hand side, only the order in which they are used does, so a producer
// warp streams them with cp.async.bulk (TMA) into shared-memory rings as fast as slots are
// released and the 8 consumer warps find every record already on chip: a level costs its
// arithmetic, not an HBM round trip per batch of columns.
//   block ring   16 KB slots, granules in order of use: forward [perm | L11^-1], [L21] per pair,
//                backward [E], [F] per pair; all 8 warps share one granule (4 columns each)
//   U ring       8.25 KB slots, one per pair of the backward sweep; up to min(8, nu) pairs of a
//                level are first reduced to their right-hand sides r = g - E z_l - F z_r (all
//                warps, pair after pair), then solved at the same time, one warp per pair
// Every consumer warp acquires and releases every granule of the block ring in order (a warp
// that skipped a use of a slot could not tell that use from the one after next by the
// barrier's phase parity); a U slot is waited on by its solver warp only, and two consecutive
// uses of it are always separated by a CTA-wide barrier.
constexpr int GRAN = SB2;             // complex entries per block-ring slot
constexpr int RING_THREADS = 288;     // warps 0-7 consume, warp 8 produces
constexpr int NCW = 8;                // consumer warps
constexpr int CW = SB / NCW;          // columns of a 32 x 32 block per warp
constexpr int BAR_CONSUMERS = 1;      // named barrier over the 256 consumer threads

struct Ring {
  cd* slots;
  uint64_t* full;
  uint64_t* empty;
  int ns;       // slots
  int stride;   // complex entries per slot
};
struct RingPos {   // next use: slot and its phase parity
  int slot;
  uint32_t phase;
};

__device__ __forceinline__ void consumer_sync() {
  asm volatile("bar.sync %0, %1;" ::"n"(BAR_CONSUMERS), "n"(NCW * 32) : "memory");
}
__device__ __forceinline__ void advance(const Ring& rg, RingPos& p) {
  if (++p.slot == rg.ns) { p.slot = 0; p.phase ^= 1u; }
}

// thread 0, before the CTA-wide barrier that precedes any use
__device__ __forceinline__ void ring_init(const Ring& rg, int consumers) {
  for (int i = 0; i < rg.ns; ++i) {
    mbar_init(&rg.full[i], 1);
    mbar_init(&rg.empty[i], consumers);
  }
}
__device__ __forceinline__ void ring_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ const cd* ring_acquire(const Ring& rg, const RingPos& p) {
  mbar_wait(&rg.full[p.slot], p.phase);
  return rg.slots + p.slot * rg.stride;
}
__device__ __forceinline__ void ring_release(const Ring& rg, RingPos& p, int lane) {
  mbar_release_slot(&rg.empty[p.slot], lane);
  advance(rg, p);
}
// producer lane: next slot of the ring <- `count` complex entries at src.  `p.phase` is the
// parity of the use being filled; `wrapped` says whether the slot has been used before.
__device__ __forceinline__ void ring_emit(const Ring& rg, RingPos& p, bool& wrapped, const cd* src,
                                          int count, uint64_t policy) {
  if (wrapped) mbar_wait(&rg.empty[p.slot], p.phase ^ 1u);
  const uint32_t bytes = static_cast<uint32_t>(count * sizeof(cd));
  mbar_expect_tx(&rg.full[p.slot], bytes);
  bulk_g2s_hint(rg.slots + p.slot * rg.stride, src, bytes, &rg.full[p.slot], policy);
  if (p.slot == rg.ns - 1) wrapped = true;
  advance(rg, p);
}

struct Producer {
  RingPos blk{0, 0u}, u{0, 0u};
  bool blk_wrapped = false, u_wrapped = false;
  uint64_t policy = 0;   // L2 eviction priority of the records this CTA streams
};

template <bool FWD>
__device__ __forceinline__ void ring_produce(const StageArgs& a, const Ring& rg, const Ring& ur,
                                             Producer& pr, int r0, int cnt) {
  for (int step = 0; step < a.mu; ++step) {
    const int lam = FWD ? step : a.mu - 1 - step;
    const int ml = (cnt + (1 << lam) - 1) >> lam;
    const size_t pair0 = a.lv[lam].off_pairs + (static_cast<size_t>(r0 >> lam) >> 1);
    for (int i = 0; i < ml / 2; ++i) {
      const cd* rec = a.pairs + (pair0 + i) * PAIR_STRIDE;
      if (FWD) {
        ring_emit(rg, pr.blk, pr.blk_wrapped, rec + PR_PERM, PR_L21 - PR_PERM, pr.policy);
        ring_emit(rg, pr.blk, pr.blk_wrapped, rec + PR_L21, SB2, pr.policy);
      } else {
        ring_emit(rg, pr.blk, pr.blk_wrapped, rec + PR_E, SB2, pr.policy);
        ring_emit(rg, pr.blk, pr.blk_wrapped, rec + PR_F, SB2, pr.policy);
        ring_emit(ur, pr.u, pr.u_wrapped, rec + PR_U, TRI, pr.policy);
      }
    }
  }
}

__device__ __forceinline__ cd shfl_cd(cd v, int src) {
  return cd{__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)};
}

// Forward sweep of one merged pair (all consumer warps).
//   s: the two stacked right-hand sides (64, shared) ; out: reduced right-hand side (32, shared)
//   g = L11^-1 (P s)_1 is kept for the back substitution ; out = (P s)_2 - M2 (P s)_1 with
//   M2 = L21 L11^-1 formed by the factorisation: both products start from (P s)_1, so the pair
//   costs one exchange of partial sums instead of two dependent ones.
//   part: 2 x NCW x 32 partial sums
__device__ __forceinline__ void pair_forward(const Ring& rg, RingPos& pos, const cd* s, cd* out,
                                             cd* __restrict__ gout, cd* part, int warp, int lane) {
  const int c0 = warp * CW;
  const cd* g0 = ring_acquire(rg, pos);
  const uint8_t* perm = reinterpret_cast<const uint8_t*>(g0 + PR_PERM);
  // the CW entries of (P s)_1 that multiply this warp's columns: one word of the permutation, then
  // broadcast reads of the right-hand side (no shuffles on the critical path of the pair)
  static_assert(CW == 4, "one 32-bit word holds the permutation entries of a warp's columns");
  const uint32_t pw = *reinterpret_cast<const uint32_t*>(perm + c0);
  cd v1[CW];
#pragma unroll
  for (int c = 0; c < CW; ++c) v1[c] = s[(pw >> (8 * c)) & 0xffu];
  // (P s)_2 of the row this lane finishes below (lanes 0 .. CW-1)
  const cd v2row = s[perm[SB + c0 + (lane & (CW - 1))]];
  const cd* L = g0 + PR_L11I;
  cd ga{0.0, 0.0}, gb{0.0, 0.0};
#pragma unroll
  for (int c = 0; c < CW; c += 2) {
    const int col = c0 + c;
    const cd ma = lane >= col ? L[tri_lo_off(col) + lane - col] : cd{0.0, 0.0};
    const cd mb = lane >= col + 1 ? L[tri_lo_off(col + 1) + lane - col - 1] : cd{0.0, 0.0};
    cfma(ga, ma, v1[c]);
    cfma(gb, mb, v1[c + 1]);
  }
  ring_release(rg, pos, lane);
  const cd* M = ring_acquire(rg, pos);
  cd aa{0.0, 0.0}, ab{0.0, 0.0};
#pragma unroll
  for (int c = 0; c < CW; c += 2) {
    cfma(aa, M[(c0 + c) * SB + lane], v1[c]);
    cfma(ab, M[(c0 + c + 1) * SB + lane], v1[c + 1]);
  }
  ring_release(rg, pos, lane);
  part[warp * SB + lane] = ga + gb;
  part[(NCW + warp) * SB + lane] = aa + ab;
  consumer_sync();
  // warp w finishes rows 4w .. 4w+3: lanes 0-3 the reduced right-hand side, lanes 4-7 g
  if (lane < 2 * CW) {
    const int row = c0 + (lane & (CW - 1));
    const cd* src = part + (lane < CW ? NCW * SB : 0) + row;
    cd t0 = src[0], t1 = src[SB];
#pragma unroll
    for (int w = 2; w < NCW; w += 2) { t0 += src[w * SB]; t1 += src[(w + 1) * SB]; }
    t0 += t1;
    if (lane < CW) out[row] = v2row - t0; else gout[row] = t0;
  }
  consumer_sync();
}

// Reduce the `cnt` rows of a chunk (first row r0 of level l0) to one row; returns the buffer
// holding it.  Rows merge pairwise, an odd last row is carried up unchanged.
__device__ __forceinline__ cd* chunk_forward(const StageArgs& a, const Ring& rg, RingPos& pos, int r0,
                                             int cnt, cd* cur, cd* nxt, cd* part) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int lam = 0; lam < a.mu; ++lam) {
    const int ml = (cnt + (1 << lam) - 1) >> lam;        // rows of the chunk at this level
    const int np = ml / 2;
    const size_t pair0 = a.lv[lam].off_pairs + (static_cast<size_t>(r0 >> lam) >> 1);
    for (int i = 0; i < np; ++i)
      pair_forward(rg, pos, cur + 2 * i * SB, nxt + i * SB, a.gvec + (pair0 + i) * SB, part, warp, lane);
    if ((ml & 1) && warp == NCW - 1) nxt[np * SB + lane] = cur[2 * np * SB + lane];
    consumer_sync();
    cd* t = cur; cur = nxt; nxt = t;
  }
  return cur;
}

__device__ __forceinline__ size_t unknown_index(const StageArgs& a, int j) {
  // level-l0 unknown j -> super-node index; the last unknown of every level is node K - 1
  return j < a.m0 ? (static_cast<size_t>(j) << a.l0) : static_cast<size_t>(a.K - 1);
}

// all-ones NaN: never produced by arithmetic (generated NaNs are canonical, propagated ones carry
// an input's payload, and this pattern is never an input)
__device__ __forceinline__ cd mbox_empty() {
  const double e = __longlong_as_double(-1LL);
  return cd{e, e};
}
__device__ __forceinline__ cd mbox_poll(const cd* p) {
  double x, y;
  do {
    asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "l"(p) : "memory");
  } while (__double_as_longlong(x) == -1LL || __double_as_longlong(y) == -1LL);
  return cd{x, y};
}
// a solved super node: the solution vector, and the mailbox when other CTAs of the launch wait for it
__device__ __forceinline__ void publish_node(const StageArgs& a, size_t node, int lane, cd v) {
  a.xv[node * SB + lane] = v;
  if (a.zbox != nullptr) {
    a.zbox_clear[node * SB + lane] = mbox_empty();
    a.zbox[node * SB + lane] = v;
  }
}

// Unit upper triangular solve by one warp: lane i holds r_i and leaves with x_i.  U is packed by
// columns with rows already divided by their diagonal entry; the diagonal slot holds 1 / U_ii.
// One column per step: broadcast x_k (shuffle), one complex FMA per lane; the next columns are
// prefetched four steps ahead.  Measured alone on the B200 (scripts/micro/trisolve.cu): 1540 cycles
// per solve (48 per unknown = shuffle + two dependent DFMA).  Variants that broadcast four unknowns
// per shuffle round and let every lane solve the 4 x 4 triangle itself were slower (1900 cycles
// branch-free with the factor in registers, 3560 with a divergent update), not faster.
__device__ __forceinline__ cd unit_upper_solve(const cd* U, cd r, int lane) {
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

// Variant for the latency-bound upper stages (one substitution at a time per SM): two columns per shuffle round,
// masked coefficients - lane i multiplies column c by U(i, c) if i < c and by zero otherwise, so the lanes of the
// block get their own unknown from the same update as the rows above (no select, no branch) - and the next block's
// entries loaded ahead.  Same subtractions in the same order per entry (bit-identical); 1390 cycles against 1540
// (scripts/micro/trisolve.cu).  The diagonal and the first block do not depend on the right-hand side: the caller
// fetches them while it waits for it (UpperPre).
struct UpperPre { cd dinv, w0, w1, u01; };
__device__ __forceinline__ void upper_blk2_load(const cd* __restrict__ U, int c0, int lane, cd& w0, cd& w1, cd& u01) {
  const cd* col0 = U + tri_up_off(c0);
  const cd* col1 = U + tri_up_off(c0 + 1);
  const cd z{0.0, 0.0};
  const cd a = col0[min(lane, c0)], b = col1[min(lane, c0 + 1)];
  w0 = lane < c0 ? a : z;
  w1 = lane < c0 + 1 ? b : z;
  u01 = col1[c0];
}
__device__ __forceinline__ UpperPre unit_upper_prefetch(const cd* __restrict__ U, int lane) {
  UpperPre p;
  p.dinv = U[tri_up_off(lane) + lane];
  upper_blk2_load(U, SB - 2, lane, p.w0, p.w1, p.u01);
  return p;
}
__device__ __forceinline__ cd unit_upper_solve2(const cd* __restrict__ U, const UpperPre& p, cd r, int lane) {
  r = r * p.dinv;
  cd w0 = p.w0, w1 = p.w1, u01 = p.u01;
#pragma unroll
  for (int kb = SB / 2 - 1; kb >= 0; --kb) {
    const int c0 = 2 * kb;
    cd n0 = w0, n1 = w1, n01 = u01;
    if (kb > 0) upper_blk2_load(U, c0 - 2, lane, n0, n1, n01);
    cd x0 = shfl_cd(r, c0);
    const cd x1 = shfl_cd(r, c0 + 1);
    cfms(r, w1, x1);
    cfms(x0, u01, x1);
    cfms(r, w0, x0);
    w0 = n0; w1 = n1; u01 = n01;
  }
  return r;
}

// z slots 0 and cnt hold the known end unknowns; fill in the interior ones.
//   z = U^-1 (g - E z_left - F z_right) per merged pair, upper levels first
__device__ __forceinline__ void chunk_backward(const StageArgs& a, const Ring& rg, RingPos& pos,
                                               const Ring& ur, RingPos& upos, int r0, int cnt, cd* z,
                                               cd* part) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = warp * CW;
  const int batch = min(NCW, ur.ns);
  int flip = 0;
  for (int lam = a.mu - 1; lam >= 0; --lam) {
    const int s = 1 << lam;
    const int ml = (cnt + s - 1) >> lam;
    const int np = ml / 2;
    const size_t pair0 = a.lv[lam].off_pairs + (static_cast<size_t>(r0 >> lam) >> 1);
    for (int i0 = 0; i0 < np; i0 += batch) {
      const int nb = min(batch, np - i0);
      const bool solver = warp < nb;
      cd r{0.0, 0.0};
      if (solver) r = a.gvec[(pair0 + i0 + warp) * SB + lane];   // in flight during the matvecs
      RingPos mine{0, 0u};
      for (int j = 0; j < nb; ++j) {
        const int ql = 2 * (i0 + j) * s, qr = min(ql + 2 * s, cnt);
        const cd* zl = z + ql * SB;
        const cd* zr = z + qr * SB;
        cd ae{0.0, 0.0}, af{0.0, 0.0};
        const cd* E = ring_acquire(rg, pos);
#pragma unroll
        for (int c = 0; c < CW; ++c) cfma(ae, E[(c0 + c) * SB + lane], zl[c0 + c]);
        ring_release(rg, pos, lane);
        const cd* F = ring_acquire(rg, pos);
#pragma unroll
        for (int c = 0; c < CW; ++c) cfma(af, F[(c0 + c) * SB + lane], zr[c0 + c]);
        ring_release(rg, pos, lane);
        cd* pj = part + flip * NCW * SB;
        flip ^= 1;
        pj[warp * SB + lane] = ae + af;
        if (j == warp) mine = upos;
        advance(ur, upos);
        consumer_sync();
        if (j == warp) {
          cd acc = pj[lane], acc2 = pj[SB + lane];
#pragma unroll
          for (int w = 2; w < NCW; w += 2) { acc += pj[w * SB + lane]; acc2 += pj[(w + 1) * SB + lane]; }
          r = r - (acc + acc2);
        }
      }
      if (solver) {
        mbar_wait(&ur.full[mine.slot], mine.phase);
        r = unit_upper_solve(ur.slots + mine.slot * ur.stride, r, lane);
        const int qm = 2 * (i0 + warp) * s + s;
        z[qm * SB + lane] = r;
        publish_node(a, unknown_index(a, r0 + qm), lane, r);
        mbar_release_slot(&ur.empty[mine.slot], lane);
      }
      consumer_sync();
    }
  }
}

// shared-memory carve-up helpers (all kernels: block ring first, 128-byte aligned)
struct SmemCursor {
  unsigned char* p;
  template <typename T>
  __device__ T* take(size_t n) {
    T* r = reinterpret_cast<T*>(p);
    p += n * sizeof(T);
    return r;
  }
};

// ---- stage bodies (consumer warps only) --------------------------------------------------
// Right-hand sides and boundary unknowns may have been written by another CTA of the same
// launch (fused upper stages): they are read past L1 (ld.global.cg).
__device__ __forceinline__ cd ldcg_cd(const cd* p) {
  const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
  return cd{v.x, v.y};
}

__device__ __forceinline__ void fwd_stage_body(const StageArgs& a, int chunk, const Ring& rg, RingPos& pos,
                                               cd* buf0, cd* buf1, cd* part) {
  const int C = 1 << a.mu;
  const int r0 = chunk * C;
  const int cnt = min(C, a.m0 - r0);
  if (a.poll_in) {
    for (int e = threadIdx.x; e < cnt * SB; e += NCW * 32) buf0[e] = mbox_poll(a.fin + static_cast<size_t>(r0) * SB + e);
  } else if (a.fin_is_rhs && a.ell.x != nullptr) {   // level-0 row k = entries (2k + 1) 16 ... of the right-hand side
    // two entries per thread and round: the value / column loads of both, then the gathers of both, are in flight
    // together (one entry at a time, the four dependent memory round trips of a thread cost the kernel 6 us)
    const size_t base = BLK + static_cast<size_t>(r0) * SB;
    for (int e0 = threadIdx.x; e0 < cnt * SB; e0 += 2 * NCW * 32) {
      const int e1 = e0 + NCW * 32;
      const bool two = e1 < cnt * SB;
      cd r0v, r1v;
      rhs_entry2(a, base + e0, base + (two ? e1 : e0), r0v, r1v);
      buf0[e0] = r0v;
      if (two) buf0[e1] = r1v;
    }
  } else {
    for (int e = threadIdx.x; e < cnt * SB; e += NCW * 32) buf0[e] = ldcg_cd(a.fin + static_cast<size_t>(r0) * SB + e);
  }
  consumer_sync();
  const cd* res = chunk_forward(a, rg, pos, r0, cnt, buf0, buf1, part);
  if (threadIdx.x < SB) {
    if (a.fout_clear != nullptr) a.fout_clear[static_cast<size_t>(chunk) * SB + threadIdx.x] = mbox_empty();
    a.fout[static_cast<size_t>(chunk) * SB + threadIdx.x] = res[threadIdx.x];
  }
}

__device__ __forceinline__ void bwd_stage_body(const StageArgs& a, int chunk, const Ring& rg, RingPos& pos,
                                               const Ring& ur, RingPos& upos, cd* z, cd* part) {
  const int C = 1 << a.mu;
  const int r0 = chunk * C;
  const int cnt = min(C, a.m0 - r0);
  if (threadIdx.x < 2 * SB) {
    const int t = threadIdx.x & (SB - 1);
    const bool right = threadIdx.x >= SB;
    const size_t idx = unknown_index(a, right ? r0 + cnt : r0) * SB + t;
    z[(right ? cnt * SB : 0) + t] = a.poll_z ? mbox_poll(a.zbox + idx) : ldcg_cd(a.xv + idx);
  }
  consumer_sync();
  chunk_backward(a, rg, pos, ur, upos, r0, cnt, z, part);
}

// stage the dense top U (64 x 64) into shared memory; waited for inside top_stage_body
__device__ __forceinline__ void top_stage_prefetch(const StageArgs& a, cd* big) {
  for (int e = threadIdx.x; e < 64 * 64; e += NCW * 32) cp_async16(big + e, a.top + TOP_U + e);
  cp_async_commit();
}

// remaining levels forward, dense top system, back substitution (one CTA)
__device__ __forceinline__ void top_stage_body(const StageArgs& a, const Ring& rg, RingPos& pos, const Ring& ur,
                                               RingPos& upos, cd* buf0, cd* buf1, cd* z, cd* part, cd* big) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cnt = a.m0;
  const int TS = a.top_size;
  if (a.poll_in) {
    for (int e = tid; e < cnt * SB; e += NCW * 32) buf0[e] = mbox_poll(a.fin + e);
  } else {
    for (int e = tid; e < cnt * SB; e += NCW * 32) buf0[e] = ldcg_cd(a.fin + e);
  }
  consumer_sync();
  const cd* res = buf0;
  if (cnt > 0) res = chunk_forward(a, rg, pos, 0, cnt, buf0, buf1, part);
  // ---- top system: [boundary row of node 0 ; last reduced row ; boundary row of node n_pad-1]
  cd* t = big + 64 * 64;
  if (tid < 64) {
    cd v{0.0, 0.0};
    if (tid < 16) v = rhs_entry(a, tid);
    else if (TS == 64 && tid < 48) v = res[tid - 16];
    else if (tid < TS) {
      const int e = tid - (TS - 16);
      if (a.n_pad - 1 < a.n) v = rhs_entry(a, static_cast<size_t>(a.n_pad - 1) * 16 + e);
    }
    t[tid] = v;
  }
  consumer_sync();
  // y = Linv (P t): 8 warps x 8 columns each, two rows per lane
  {
    const uint8_t* perm = reinterpret_cast<const uint8_t*>(a.top);
    const cd* Linv = a.top + TOP_LINV;
    cd m0[8], m1[8], p0{0.0, 0.0}, p1{0.0, 0.0};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      m0[c] = ldg_cd(Linv + (warp * 8 + c) * 64 + lane);
      m1[c] = ldg_cd(Linv + (warp * 8 + c) * 64 + 32 + lane);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const cd tv = t[perm[warp * 8 + c]];
      cfma(p0, m0[c], tv);
      cfma(p1, m1[c], tv);
    }
    part[warp * 64 + lane] = p0;
    part[warp * 64 + 32 + lane] = p1;
  }
  cp_async_wait_all();
  consumer_sync();
  if (warp == 0) {
    cd y0{0.0, 0.0}, y1{0.0, 0.0};   // rows lane and lane + 32
#pragma unroll
    for (int w = 0; w < 8; ++w) { y0 += part[w * 64 + lane]; y1 += part[w * 64 + 32 + lane]; }
    // unit upper triangular back substitution (rows pre-divided by their diagonal, whose
    // reciprocal sits on the diagonal), column-major with leading dimension 64
    y0 = y0 * big[lane * 64 + lane];
    y1 = y1 * big[(lane + 32) * 64 + lane + 32];
    for (int k = TS - 1; k >= 1; --k) {
      const cd* col = big + k * 64;
      const cd xk = shfl_cd(k >= 32 ? y1 : y0, k & 31);
      if (k >= 32) {
        if (lane + 32 < k) cfms(y1, col[lane + 32], xk);
        cfms(y0, col[lane], xk);
      } else if (lane < k) {
        cfms(y0, col[lane], xk);
      }
    }
    z[lane] = y0;
    publish_node(a, 0, lane, y0);
    if (TS == 64) {
      z[cnt * SB + lane] = y1;
      publish_node(a, static_cast<size_t>(a.K - 1), lane, y1);
    }
  }
  consumer_sync();
  if (cnt > 0) chunk_backward(a, rg, pos, ur, upos, 0, cnt, z, part);
}

// smem: [ns slots][buf0 C x 32][buf1 C x 32][part 2 x NCW x 32][2 ns barriers]
__global__ void __launch_bounds__(RING_THREADS, 3) slu_fwd_stage_kernel(StageArgs a, int ns) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int C = 1 << a.mu;
  SmemCursor sc{smem_raw};
  cd* slots = sc.take<cd>(static_cast<size_t>(ns) * GRAN);
  cd* buf0 = sc.take<cd>(C * SB);
  cd* buf1 = sc.take<cd>(C * SB);
  cd* part = sc.take<cd>(2 * NCW * SB);
  uint64_t* bars = sc.take<uint64_t>(2 * ns);
  const Ring rg{slots, bars, bars + ns, ns, GRAN};
  TRACE_MARK(0, 0);
  TRACE_VALUE(0, 15, trace_smid());
  if (threadIdx.x == 0) { ring_init(rg, NCW); ring_init_fence(); }
  __syncthreads();
  pdl_launch_dependents();
  if (threadIdx.x >= NCW * 32) {
    if (threadIdx.x == NCW * 32) {   // the factor records are static: streamed before the dependency resolves
      Producer pr;
      pr.policy = gridDim.x > 148 ? l2_policy_evict_first() : l2_policy_evict_last();
      const int r0 = blockIdx.x * C;
      ring_produce<true>(a, rg, rg, pr, r0, min(C, a.m0 - r0));
    }
    return;
  }
  pdl_wait();
  TRACE_MARK(0, 1);
  RingPos pos{0, 0u};
  fwd_stage_body(a, blockIdx.x, rg, pos, buf0, buf1, part);
  TRACE_MARK(0, 2);
}

// smem: [ns slots][nu U slots][z (C + 1) x 32][part][barriers]
__global__ void __launch_bounds__(RING_THREADS, 3) slu_bwd_stage_kernel(StageArgs a, int ns, int nu) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int C = 1 << a.mu;
  SmemCursor sc{smem_raw};
  cd* slots = sc.take<cd>(static_cast<size_t>(ns) * GRAN);
  cd* uslots = sc.take<cd>(static_cast<size_t>(nu) * TRI);
  cd* z = sc.take<cd>((C + 1) * SB);
  cd* part = sc.take<cd>(2 * NCW * SB);
  uint64_t* bars = sc.take<uint64_t>(2 * (ns + nu));
  const Ring rg{slots, bars, bars + ns, ns, GRAN};
  const Ring ur{uslots, bars + 2 * ns, bars + 2 * ns + nu, nu, TRI};
  TRACE_MARK(2, 0);
  TRACE_VALUE(2, 15, trace_smid());
  if (threadIdx.x == 0) { ring_init(rg, NCW); ring_init(ur, 1); ring_init_fence(); }
  __syncthreads();
  pdl_launch_dependents();
  if (threadIdx.x >= NCW * 32) {
    if (threadIdx.x == NCW * 32) {
      Producer pr;
      pr.policy = gridDim.x > 148 ? l2_policy_evict_first() : l2_policy_evict_last();
      const int r0 = blockIdx.x * C;
      ring_produce<false>(a, rg, ur, pr, r0, min(C, a.m0 - r0));
    }
    return;
  }
  pdl_wait();
  TRACE_MARK(2, 1);
  RingPos pos{0, 0u}, upos{0, 0u};
  bwd_stage_body(a, blockIdx.x, rg, pos, ur, upos, z, part);
  TRACE_MARK(2, 2);
  if (a.done_flags != nullptr) {   // this chunk's part of the solution is final (SolveSignal)
    consumer_sync();
    if (threadIdx.x == 0) {
      __threadfence();
      asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(a.done_flags + blockIdx.x), "l"(a.done_epoch) : "memory");
    }
  }
}

// The narrow upper stages and the top system in ONE cooperative launch (grid = chunks of the first
// fused stage, all co-resident): their data is a few MB, their cost is the dependency chain,
// and a kernel boundary per stage (launch gap, ring start-up) would double it.  CTAs hand rows up
// and unknowns down through mailboxes (SluDevice::mbox): a consumer polls the VALUE it needs until
// it is no longer the all-ones NaN, so a hand-off costs one store and one load round trip and a
// chunk waits for its own inputs only (the first version counted finished chunks of a whole
// stage with fence + atomic on one side and poll + barrier + load on the other: ~2.5 k cycles per
// hand-off, eight of them on the critical path).  The producer warp never waits (the factor
// records are static), so the next stage's records are on chip when its dependencies arrive.
struct FusedArgs {
  int nst;                       // fused stages; the last one is the top stage
  int ns, nu, cmax;              // ring shapes, rows of the widest chunk
  StageArgs st[MAX_FUSED];
};

// smem: [ns slots][nu U slots][buf0][buf1][z][part][big 64 x 64 + 64][barriers], sized for cmax
__global__ void __launch_bounds__(RING_THREADS) slu_fused_stage_kernel(const __grid_constant__ FusedArgs f) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, b = blockIdx.x;
  SmemCursor sc{smem_raw};
  cd* slots = sc.take<cd>(static_cast<size_t>(f.ns) * GRAN);
  cd* uslots = sc.take<cd>(static_cast<size_t>(f.nu) * TRI);
  cd* buf0 = sc.take<cd>(f.cmax * SB);
  cd* buf1 = sc.take<cd>(f.cmax * SB);
  cd* z = sc.take<cd>((f.cmax + 1) * SB);
  cd* part = sc.take<cd>(2 * NCW * SB);
  cd* big = sc.take<cd>(64 * 64 + 64);
  uint64_t* bars = sc.take<uint64_t>(2 * (f.ns + f.nu));
  const Ring rg{slots, bars, bars + f.ns, f.ns, GRAN};
  const Ring ur{uslots, bars + 2 * f.ns, bars + 2 * f.ns + f.nu, f.nu, TRI};
  if (tid == 0) { ring_init(rg, NCW); ring_init(ur, 1); ring_init_fence(); }
  __syncthreads();
  const int top = f.nst - 1;
  if (tid >= NCW * 32) {
    if (tid == NCW * 32) {
      Producer pr;
      pr.policy = l2_policy_evict_last();   // the upper levels: a few MB that every solve re-reads
      for (int s = 0; s < top; ++s) {
        const StageArgs& a = f.st[s];
        const int C = 1 << a.mu;
        if (b < a.nchunks) ring_produce<true>(a, rg, ur, pr, b * C, min(C, a.m0 - b * C));
      }
      if (b == 0 && f.st[top].m0 > 0) {
        ring_produce<true>(f.st[top], rg, ur, pr, 0, f.st[top].m0);
        ring_produce<false>(f.st[top], rg, ur, pr, 0, f.st[top].m0);
      }
      for (int s = top - 1; s >= 0; --s) {
        const StageArgs& a = f.st[s];
        const int C = 1 << a.mu;
        if (b < a.nchunks) ring_produce<false>(a, rg, ur, pr, b * C, min(C, a.m0 - b * C));
      }
    }
    return;
  }
  if (b == 0) top_stage_prefetch(f.st[top], big);
  RingPos pos{0, 0u}, upos{0, 0u};
  for (int s = 0; s < top; ++s) {
    const StageArgs& a = f.st[s];
    if (b >= a.nchunks) continue;
    fwd_stage_body(a, b, rg, pos, buf0, buf1, part);   // inputs of stages > 0: polled from the mailbox
    consumer_sync();                                   // buf0 / buf1 are reused by the next stage
  }
  if (b == 0) {
    top_stage_body(f.st[top], rg, pos, ur, upos, buf0, buf1, z, part, big);
    consumer_sync();
  }
  for (int s = top - 1; s >= 0; --s) {
    const StageArgs& a = f.st[s];
    if (b >= a.nchunks) continue;
    bwd_stage_body(a, b, rg, pos, ur, upos, z, part);  // boundary unknowns: polled from the mailbox
    consumer_sync();
  }
}

// ---- upper stages with RESIDENT factor records -----------------------------------------------
// The levels above the first stage hold 312 pairs at G = 10 001 (20 MB of records): their cost is
// not bytes but a chain of ~2 x 9 dependent pair steps plus the dense top system.  The version
// above streams the records of a chunk through rings and lets CTA 0 take part in every stage, so
// every pair step also waits for its ~66 KB to pass the SM's copy engine (>= 0.6 us per pair) and
// runs all 8 warps + two CTA barriers per pair.  Here instead
//   * every chunk of every stage and the top system get a CTA (SM) of their own: 79 + 20 + 5 + 2 + 1
//     at G = 10 001, all co-resident; a chunk has at most three pairs (mu <= 2), whose records
//     (<= 200 KB) are copied into shared memory ONCE, by bulk copies issued in the prologue - before
//     griddepcontrol.wait, i.e. while the previous kernel of the solve is still running when the
//     launch is programmatic - so no pair step ever waits for a copy;
//   * a pair is worked on by four warps (thread = row r of an octet, column group q: 8 columns
//     each, two shuffle rounds instead of a partial-sum exchange through shared memory), so the
//     two pairs of a level run concurrently on the two halves of the CTA;
// Hand-offs between CTAs are the same mailboxes as above.
constexpr int UP_THREADS = 256;
constexpr int UP_MAXPAIRS = 3;                              // mu <= 2: two pairs + one pair
constexpr int UP_REC_ELEMS = UP_MAXPAIRS * PAIR_STRIDE;     // also fits one pair + the dense top record
static_assert(PAIR_STRIDE + TOP_STRIDE <= UP_REC_ELEMS, "top CTA: one pair record + the dense top record");
constexpr int UP_FWD_ELEMS = PR_E - PR_PERM;                // [perm | L11^-1 | M2]
constexpr int UP_BWD_ELEMS = PR_U + TRI - PR_E;             // [E | F | U]

struct UpperArgs {
  int nst;                       // stages of this launch; the last one is the top stage
  int cta0[MAX_FUSED + 1];       // first CTA of each stage
  StageArgs st[MAX_FUSED];
  // Back-substitution records [E | F | U] of the first stage's upper levels: pf_count records from pf_first on.
  // This launch is a dependency chain that leaves HBM idle; the CTAs of its first stage use the wait for their
  // boundary unknowns to pull those records into L2 (cp.async.bulk.prefetch.L2), so that the backward first-stage
  // kernel that follows streams only its lowest level from HBM.
  const cd* pf_first;
  int pf_count;
};

__device__ __forceinline__ void l2_prefetch(const void* src, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(bytes), "l"(policy) : "memory");
}

__device__ __forceinline__ void group_sync(int grp) {
  asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
}
__device__ __forceinline__ cd quad_sum(cd v) {   // over the four column groups: lanes l, l ^ 8, l ^ 16, l ^ 24
  v.x += __shfl_xor_sync(0xffffffffu, v.x, 8);
  v.y += __shfl_xor_sync(0xffffffffu, v.y, 8);
  v.x += __shfl_xor_sync(0xffffffffu, v.x, 16);
  v.y += __shfl_xor_sync(0xffffffffu, v.y, 16);
  return v;
}

// One pair is worked on by 128 threads: thread = (row r of an octet of rows, column group q = 8 columns q, q + 4, ...).
// The factor records are static, so every thread copies its 16 + 16 matrix entries of the forward step (and later
// of the back substitution) from shared memory into REGISTERS while it waits for its inputs; a pair step on the
// critical path is then 8 broadcast reads of the right-hand side, 16 complex FMAs and two shuffle rounds, instead
// of 24 - 32 KB streamed from shared memory per pair (measured with the kernel timeline: 0.75 us per forward level
// and 0.32 us per right-hand side before).
struct FwdRegs {
  cd m[8], l[8];   // M2(row, c_k) and L11^-1(row, c_k) (zero above the diagonal), c_k = q + 4 k
  int src[8];      // perm[c_k]: which entry of the stacked right-hand sides multiplies column c_k
  int src2;        // perm[32 + row]
};
struct BwdRegs {
  cd e[8], f[8];   // E(row, c_k), F(row, c_k)
};
__device__ __forceinline__ void load_fwd(const cd* rec, int t, FwdRegs& R) {
  const int lane = t & 31, q = lane >> 3, row = 8 * (t >> 5) + (lane & 7);
  const uint8_t* perm = reinterpret_cast<const uint8_t*>(rec + PR_PERM);
  const cd* L = rec + PR_L11I;
  const cd* M = rec + PR_L21;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = q + 4 * k;
    R.m[k] = M[c * SB + row];
    R.l[k] = c <= row ? L[tri_lo_off(c) + row - c] : cd{0.0, 0.0};
    R.src[k] = perm[c];
  }
  R.src2 = perm[SB + row];
}
__device__ __forceinline__ void load_bwd(const cd* rec, int t, BwdRegs& R) {
  const int lane = t & 31, q = lane >> 3, row = 8 * (t >> 5) + (lane & 7);
  const cd* E = rec + PR_E;
  const cd* F = rec + PR_F;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    R.e[k] = E[(q + 4 * k) * SB + row];
    R.f[k] = F[(q + 4 * k) * SB + row];
  }
}
// forward sweep of one pair: out = (P s)_2 - M2 (P s)_1 ; returns g = L11^-1 (P s)_1 of this thread's row (q == 0)
__device__ __forceinline__ cd up_pair_forward(const FwdRegs& R, const cd* s, cd* out, int t) {
  const int lane = t & 31, q = lane >> 3, row = 8 * (t >> 5) + (lane & 7);
  cd v1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v1[k] = s[R.src[k]];
  const cd s2 = s[R.src2];
  cd ga{0.0, 0.0}, gb{0.0, 0.0}, oa{0.0, 0.0}, ob{0.0, 0.0};
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    cfma(oa, R.m[k], v1[k]);
    cfma(ob, R.m[k + 1], v1[k + 1]);
    cfma(ga, R.l[k], v1[k]);
    cfma(gb, R.l[k + 1], v1[k + 1]);
  }
  const cd gs = quad_sum(ga + gb), os = quad_sum(oa + ob);
  if (q == 0) out[row] = s2 - os;
  return gs;
}
// right-hand side of the back substitution of one pair: r = g - E z_left - F z_right
__device__ __forceinline__ void up_pair_backward_rhs(const BwdRegs& R, const cd* zl, const cd* zr, cd g, cd* r, int t) {
  const int lane = t & 31, q = lane >> 3, row = 8 * (t >> 5) + (lane & 7);
  cd ea{0.0, 0.0}, eb{0.0, 0.0}, fa{0.0, 0.0}, fb{0.0, 0.0};
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    const int c = q + 4 * k, d = c + 4;
    cfma(ea, R.e[k], zl[c]);
    cfma(eb, R.e[k + 1], zl[d]);
    cfma(fa, R.f[k], zr[c]);
    cfma(fb, R.f[k + 1], zr[d]);
  }
  const cd acc = quad_sum((ea + eb) + (fa + fb));
  if (q == 0) r[row] = g - acc;
}

// 64 x 64 unit upper triangular solve by one warp (rows lane and lane + 32), column-major with leading
// dimension 64, diagonal slot = 1 / U_ii: one column per step like unit_upper_solve (shuffle + one complex FMA per
// row), the loop unrolled by 8 only.  History (scripts/micro/trisolve.cu and the kernel timeline): four columns per
// shuffle round with a divergent update 8950 cycles; branch-free and fully unrolled 3650 cycles warm, but 64 KB of
// straight-line code that runs once per launch - 33 us with a cold instruction cache; the same as rolled loops
// 3.7 us in the kernel.
__device__ __forceinline__ void unit_upper_solve64(const cd* __restrict__ U, cd& y0, cd& y1, int lane) {
  y0 = y0 * U[lane * 64 + lane];
  y1 = y1 * U[(lane + 32) * 64 + lane + 32];
#pragma unroll 8
  for (int k = 63; k >= 32; --k) {
    const cd* col = U + k * 64;
    const cd u0 = col[lane], u1 = col[min(lane + 32, k)];
    const cd xk = shfl_cd(y1, k - 32);
    cfms(y0, u0, xk);
    if (lane + 32 < k) cfms(y1, u1, xk);
  }
#pragma unroll 8
  for (int k = 31; k >= 1; --k) {
    const cd u0 = U[k * 64 + min(lane, k)];
    const cd xk = shfl_cd(y0, k);
    if (lane < k) cfms(y0, u0, xk);
  }
}

// smem: [3 pair records | (top CTA) 1 pair record + dense top record][level rows 2 x 4 x 32]
//       [z 5 x 32][r 2 x 32][t 64][partial sums 4 x 64][7 barriers]
constexpr size_t UP_SMEM = sizeof(cd) * (UP_REC_ELEMS + 2 * 4 * SB + 5 * SB + 2 * SB + 64 + 4 * 64) +
                           sizeof(uint64_t) * (2 * UP_MAXPAIRS + 2);

__global__ void __launch_bounds__(UP_THREADS, 1) slu_upper_kernel(const __grid_constant__ UpperArgs f) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemCursor sc{smem_raw};
  cd* recs = sc.take<cd>(UP_REC_ELEMS);
  cd* rows = sc.take<cd>(2 * 4 * SB);
  cd* z = sc.take<cd>(5 * SB);
  cd* rbuf = sc.take<cd>(2 * SB);
  cd* tvec = sc.take<cd>(64);
  cd* part = sc.take<cd>(4 * 64);
  uint64_t* bars = sc.take<uint64_t>(2 * UP_MAXPAIRS + 2);   // [2p] forward part of pair p, [2p + 1] backward part, [6] top record
  const int tid = threadIdx.x, lane = tid & 31, b = blockIdx.x;
  const int grp = tid >> 7, t = tid & 127;
  TRACE_MARK(1, 0);
  int s = 0;
  while (s + 1 < f.nst && b >= f.cta0[s + 1]) ++s;
  TRACE_VALUE(1, 15, (static_cast<unsigned long long>(s) << 32) | trace_smid());
  const StageArgs& a = f.st[s];
  const bool top = s == f.nst - 1;
  const int chunk = b - f.cta0[s];
  const int mu = a.mu;
  const int r0 = top ? 0 : chunk << mu;
  const int cnt = top ? a.m0 : min(1 << mu, a.m0 - r0);
  // pairs of the chunk: np0 <= 2 at its first level (record slots 0, 1), np1 <= 1 at its second (slot np0)
  const int np0 = mu >= 1 ? cnt / 2 : 0;
  const int m1 = (cnt + 1) >> 1;
  const int np1 = mu >= 2 ? m1 / 2 : 0;
  // ---- prologue: every record this CTA will ever use -> shared memory (static data: no dependence on
  // the kernels before this one), in order of use
  if (tid == 0) {
    for (int i = 0; i < 2 * UP_MAXPAIRS + 1; ++i) mbar_init(&bars[i], 1);
    ring_init_fence();
    const uint64_t keep = l2_policy_evict_last();
    const cd* rec0 = a.pairs + (a.lv[0].off_pairs + (static_cast<size_t>(r0) >> 1)) * PAIR_STRIDE;
    const cd* rec1 = mu >= 2 ? a.pairs + (a.lv[1].off_pairs + (static_cast<size_t>(r0 >> 1) >> 1)) * PAIR_STRIDE : nullptr;
    constexpr uint32_t fwd_bytes = sizeof(cd) * UP_FWD_ELEMS, bwd_bytes = sizeof(cd) * UP_BWD_ELEMS;
    for (int i = 0; i < np0; ++i) {
      mbar_expect_tx(&bars[2 * i], fwd_bytes);
      bulk_g2s_hint(recs + i * PAIR_STRIDE + PR_PERM, rec0 + i * PAIR_STRIDE + PR_PERM, fwd_bytes, &bars[2 * i], keep);
    }
    if (np1) {
      mbar_expect_tx(&bars[2 * np0], fwd_bytes);
      bulk_g2s_hint(recs + np0 * PAIR_STRIDE + PR_PERM, rec1 + PR_PERM, fwd_bytes, &bars[2 * np0], keep);
    }
    if (top) {
      constexpr uint32_t bytes = sizeof(cd) * TOP_STRIDE;
      mbar_expect_tx(&bars[2 * UP_MAXPAIRS], bytes);
      bulk_g2s_hint(recs + PAIR_STRIDE, a.top, bytes, &bars[2 * UP_MAXPAIRS], keep);
    }
    if (np1) {
      mbar_expect_tx(&bars[2 * np0 + 1], bwd_bytes);
      bulk_g2s_hint(recs + np0 * PAIR_STRIDE + PR_E, rec1 + PR_E, bwd_bytes, &bars[2 * np0 + 1], keep);
    }
    for (int i = 0; i < np0; ++i) {
      mbar_expect_tx(&bars[2 * i + 1], bwd_bytes);
      bulk_g2s_hint(recs + i * PAIR_STRIDE + PR_E, rec0 + i * PAIR_STRIDE + PR_E, bwd_bytes, &bars[2 * i + 1], keep);
    }
  }
  __syncthreads();
  TRACE_MARK(1, 1);
  pdl_launch_dependents();
  // ---- this thread's share of the forward matrices -> registers.  Group g owns pair g of the first level (A);
  // group 0 also owns the pair of the second level (B).
  const bool hasA = grp < np0, hasB = grp == 0 && np1 > 0;
  FwdRegs FA, FB;
  BwdRegs BA, BB;
  UpperPre preA, preB;
  if (hasA) { mbar_wait(&bars[2 * grp], 0u); load_fwd(recs + grp * PAIR_STRIDE, t, FA); }
  if (hasB) { mbar_wait(&bars[2 * np0], 0u); load_fwd(recs + np0 * PAIR_STRIDE, t, FB); }
  // dense top system: thread = (row, 16 columns) of Linv; the top CTA has no second-level pair, so the entries
  // live in the registers of FB (m: columns 0 .. 7 of the thread's 16, l: 8 .. 15, src[0 .. 3]: permutation bytes)
  const cd* toprec = recs + PAIR_STRIDE;
  if (top) {
    mbar_wait(&bars[2 * UP_MAXPAIRS], 0u);
    const int row = tid & 63, cg = tid >> 6;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      FB.m[c] = toprec[TOP_LINV + (cg * 16 + c) * 64 + row];
      FB.l[c] = toprec[TOP_LINV + (cg * 16 + 8 + c) * 64 + row];
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) FB.src[w] = static_cast<int>(reinterpret_cast<const uint32_t*>(toprec)[cg * 4 + w]);
  }
  pdl_wait();   // everything below reads what earlier kernels wrote
  TRACE_MARK(1, 2);
  // the top system's boundary right-hand sides do not depend on this launch: their load overlaps the wait for the rows
  cd bvec{0.0, 0.0};
  if (top && tid < 64 && (tid < 16 || tid >= 48)) {
    if (tid < 16) bvec = rhs_entry(a, tid);
    else if (a.n_pad - 1 < a.n) bvec = rhs_entry(a, static_cast<size_t>(a.n_pad - 1) * 16 + tid - 48);
  }
  // ---- input rows
  if (tid < cnt * SB) {
    const cd* src = a.fin + static_cast<size_t>(r0) * SB + tid;
    rows[tid] = a.poll_in ? mbox_poll(src) : ldcg_cd(src);
  }
  __syncthreads();
  TRACE_MARK(1, 3);
  cd* cur = rows;
  cd* nxt = rows + 4 * SB;
  // ---- forward: the pairs of a level side by side; g of a pair stays in the registers of its q == 0 threads
  cd gA{0.0, 0.0}, gB{0.0, 0.0};
  if (mu >= 1) {
    if (hasA) gA = up_pair_forward(FA, cur + 2 * grp * SB, nxt + grp * SB, t);
    if ((cnt & 1) && tid >= UP_THREADS - 32) nxt[np0 * SB + lane] = cur[2 * np0 * SB + lane];   // odd row: carried up
    __syncthreads();
    cd* tmp = cur; cur = nxt; nxt = tmp;
  }
  if (mu >= 2) {
    if (hasB) gB = up_pair_forward(FB, cur, nxt, t);
    if ((m1 & 1) && tid >= UP_THREADS - 32) nxt[np1 * SB + lane] = cur[2 * np1 * SB + lane];
    __syncthreads();
    cd* tmp = cur; cur = nxt; nxt = tmp;
  }
  TRACE_MARK(1, 4);
  if (!top) {
    if (tid < SB) {
      if (a.fout_clear != nullptr) a.fout_clear[static_cast<size_t>(chunk) * SB + tid] = mbox_empty();
      a.fout[static_cast<size_t>(chunk) * SB + tid] = cur[tid];
    }
    if (s == 0 && f.pf_count > 0 && tid == UP_THREADS - 1) {   // this CTA's share of the L2 prefetch (see UpperArgs)
      const uint64_t keep = l2_policy_evict_last();
      const int nc = f.cta0[1] - f.cta0[0];
      for (int i = chunk; i < f.pf_count; i += nc)
        l2_prefetch(f.pf_first + static_cast<size_t>(i) * PAIR_STRIDE + PR_E, sizeof(cd) * UP_BWD_ELEMS, keep);
    }
    // the matrices of the back substitution take the place of the forward ones while the unknowns are on their way
    if (hasA) { mbar_wait(&bars[2 * grp + 1], 0u); load_bwd(recs + grp * PAIR_STRIDE, t, BA); }
    if (hasB) { mbar_wait(&bars[2 * np0 + 1], 0u); load_bwd(recs + np0 * PAIR_STRIDE, t, BB); }
    if (t < 32) {   // the solver warp of the group: what the substitutions need before their right-hand sides exist
      if (hasA) preA = unit_upper_prefetch(recs + grp * PAIR_STRIDE + PR_U, lane);
      if (hasB) preB = unit_upper_prefetch(recs + np0 * PAIR_STRIDE + PR_U, lane);
    }
    // boundary unknowns of the chunk
    if (tid < 2 * SB) {
      const bool right = tid >= SB;
      const size_t idx = unknown_index(a, right ? r0 + cnt : r0) * SB + lane;
      z[(right ? cnt * SB : 0) + lane] = a.poll_z ? mbox_poll(a.zbox + idx) : ldcg_cd(a.xv + idx);
    }
  } else {
    // ---- dense top system [boundary row of node 0 ; last reduced row ; boundary row of node n_pad - 1]
    if (tid < 64) tvec[tid] = (tid >= 16 && tid < 48) ? cur[tid - 16] : bvec;
    __syncthreads();
    {   // y = Linv (P t)
      const int row = tid & 63, cg = tid >> 6;
      cd p0{0.0, 0.0}, p1{0.0, 0.0};
#pragma unroll
      for (int c = 0; c < 16; c += 2) {
        const uint32_t w = static_cast<uint32_t>(FB.src[c >> 2]);
        cfma(p0, c < 8 ? FB.m[c] : FB.l[c - 8], tvec[(w >> (8 * (c & 3))) & 0xffu]);
        cfma(p1, c < 8 ? FB.m[c + 1] : FB.l[c - 7], tvec[(w >> (8 * ((c + 1) & 3))) & 0xffu]);
      }
      part[cg * 64 + row] = p0 + p1;
    }
    if (hasA) {
      mbar_wait(&bars[2 * grp + 1], 0u);
      load_bwd(recs + grp * PAIR_STRIDE, t, BA);
      if (t < 32) preA = unit_upper_prefetch(recs + grp * PAIR_STRIDE + PR_U, lane);
    }
    __syncthreads();
    TRACE_MARK(1, 13);
    if (tid < 32) {
      cd y0 = (part[lane] + part[64 + lane]) + (part[128 + lane] + part[192 + lane]);
      cd y1 = (part[32 + lane] + part[96 + lane]) + (part[160 + lane] + part[224 + lane]);
      unit_upper_solve64(toprec + TOP_U, y0, y1, lane);
      z[lane] = y0;
      publish_node(a, 0, lane, y0);
      z[cnt * SB + lane] = y1;
      publish_node(a, static_cast<size_t>(a.K - 1), lane, y1);
    }
    TRACE_MARK(1, 14);
  }
  __syncthreads();
  TRACE_MARK(1, 5);
  // ---- backward: z = U^-1 (g - E z_left - F z_right), upper level first, pairs of a level side by side
  if (hasB) {   // second level: one pair, rows 0 .. 3 of the chunk
    const int qr = min(4, cnt);
    TRACE_MARK(1, 10);
    up_pair_backward_rhs(BB, z, z + qr * SB, gB, rbuf, t);
    group_sync(0);
    TRACE_MARK(1, 11);
    if (t < 32) {
      const cd x = unit_upper_solve2(recs + np0 * PAIR_STRIDE + PR_U, preB, rbuf[lane], lane);
      z[2 * SB + lane] = x;
      publish_node(a, unknown_index(a, r0 + 2), lane, x);
    }
    TRACE_MARK(1, 12);
  }
  if (mu >= 2) {
    __syncthreads();
    TRACE_MARK(1, 6);
  }
  if (hasA) {
    const int ql = 2 * grp, qr = min(ql + 2, cnt), qm = ql + 1;
    up_pair_backward_rhs(BA, z + ql * SB, z + qr * SB, gA, rbuf + grp * SB, t);
    group_sync(grp);
    if (t < 32) {
      const cd x = unit_upper_solve2(recs + grp * PAIR_STRIDE + PR_U, preA, rbuf[grp * SB + lane], lane);
      z[qm * SB + lane] = x;
      publish_node(a, unknown_index(a, r0 + qm), lane, x);
    }
  }
  TRACE_MARK(1, 7);
}

// Single CTA: the top stage alone (problems too small for more than one stage above the first).
// smem: [ns slots][nu U slots][buf0][buf1][z][part][big 64 x 64 + 64][barriers]
__global__ void __launch_bounds__(RING_THREADS) slu_top_stage_kernel(StageArgs a, int ns, int nu) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int C = 1 << a.mu;
  const int tid = threadIdx.x;
  SmemCursor sc{smem_raw};
  cd* slots = sc.take<cd>(static_cast<size_t>(ns) * GRAN);
  cd* uslots = sc.take<cd>(static_cast<size_t>(nu) * TRI);
  cd* buf0 = sc.take<cd>(C * SB);
  cd* buf1 = sc.take<cd>(C * SB);
  cd* z = sc.take<cd>((C + 1) * SB);
  cd* part = sc.take<cd>(2 * NCW * SB);     // doubles as the 8 x 64 scratch of the dense step
  cd* big = sc.take<cd>(64 * 64 + 64);      // staged top U, then the 64 right-hand-side entries
  uint64_t* bars = sc.take<uint64_t>(2 * (ns + nu));
  const Ring rg{slots, bars, bars + ns, ns, GRAN};
  const Ring ur{uslots, bars + 2 * ns, bars + 2 * ns + nu, nu, TRI};
  if (tid == 0) { ring_init(rg, NCW); ring_init(ur, 1); ring_init_fence(); }
  __syncthreads();
  if (tid >= NCW * 32) {
    if (tid == NCW * 32 && a.m0 > 0) {
      Producer pr;
      pr.policy = l2_policy_evict_last();
      ring_produce<true>(a, rg, ur, pr, 0, a.m0);
      ring_produce<false>(a, rg, ur, pr, 0, a.m0);
    }
    return;
  }
  top_stage_prefetch(a, big);
  RingPos pos{0, 0u}, upos{0, 0u};
  top_stage_body(a, rg, pos, ur, upos, buf0, buf1, z, part, big);
}

// ------------------------------------------------------------------ factorisation kernels
struct FactorArgs {
  const cd* A;
  const cd* B;
  cd sigma;
  int n, n_pad, K;
  const cd* src;       // work rows of this level
  cd* dst;             // work rows of the next level
  cd* pairs;           // pair records of this level
  cd* top;
  int m;               // rows at this level
  int npairs;
  int level;
  int top_size;
  int32_t* info;
  uint32_t padmask;
};

// entry (i, j) of tile t (0 sub, 1 diag, 2 super) of block row r of A - sigma*B; the padding
// node (r >= n) is a decoupled identity row
__device__ __forceinline__ cd m_entry(const FactorArgs& a, int r, int t, int i, int j) {
  if (r >= a.n) return cd{(t == 1 && i == j) ? 1.0 : 0.0, 0.0};
  const size_t off = (static_cast<size_t>(r) * 3 + t) * BLK2 + j * BLK + i;
  if (t == 1 && i == j && ((a.padmask >> i) & 1u)) return cd{1.0, 0.0};   // absent variable (hd / hd-1d)
  return a.A[off] - a.sigma * a.B[off];
}

// level-0 rows of the bidiagonal form: row k = block rows (2k+1, 2k+2) on z_k (S) and z_k+1 (T)
__global__ void __launch_bounds__(256) slu_build_rows_kernel(FactorArgs a) {
  const int k = blockIdx.x;
  const int r1 = 2 * k + 1, r2 = 2 * k + 2;
  cd* S = a.dst + static_cast<size_t>(k) * ROW_STRIDE;
  cd* T = S + SB2;
  for (int e = threadIdx.x; e < SB2; e += blockDim.x) {
    const int row = e & 31, col = e >> 5;
    const int i = row & 15, j = col & 15;
    const bool top = row < 16, left = col < 16;
    cd s{0.0, 0.0}, t{0.0, 0.0};
    if (top) {
      s = m_entry(a, r1, left ? 0 : 1, i, j);
      if (left) t = m_entry(a, r1, 2, i, j);
    } else {
      if (!left) s = m_entry(a, r2, 0, i, j);
      t = m_entry(a, r2, left ? 1 : 2, i, j);
    }
    S[e] = s;
    T[e] = t;
  }
}

constexpr int WLD = 97;   // padded row length of the 64 x 96 working matrix
constexpr int NBF = 4;    // elimination steps per pass over the panel (divides 32)

// Merge rows (2p, 2p+1) of a level: GE with partial pivoting over the 64 stacked rows of the
// panel [T_2p ; S_2p+1]; block npairs (if present) carries the odd last row up unchanged.
template <bool LOOKAHEAD>
__global__ void __launch_bounds__(256) slu_merge_kernel(FactorArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cd* W = reinterpret_cast<cd*>(smem_raw);          // [64][WLD]
  cd* lcol = W + 64 * WLD;                          // [2][NBF][64] multipliers of the columns of a block of steps (two blocks in flight)
  // inverse of L11: lives in rows 32..63, columns 32..63 of W once the reduced rows stored there
  // have been written out (keeps the CTA at 100 KB of shared memory: two CTAs per SM)
  cd* X = W + 32 * WLD + 32;                        // X(i, c) at X[i * WLD + c]
  int* prm = reinterpret_cast<int*>(lcol + 2 * NBF * 64);   // [64]
  __shared__ double rscale[64];
  __shared__ int pivs[2][NBF];                      // look-ahead: row swaps of a block not yet applied outside its panel
  const int tid = threadIdx.x;
  const int p = blockIdx.x;
  if (p == a.npairs) {   // odd row out
    const cd* s = a.src + static_cast<size_t>(a.m - 1) * ROW_STRIDE;
    cd* d = a.dst + static_cast<size_t>(a.npairs) * ROW_STRIDE;
    for (int e = tid; e < ROW_STRIDE; e += blockDim.x) d[e] = s[e];
    return;
  }
  const cd* Sa = a.src + static_cast<size_t>(2 * p) * ROW_STRIDE;
  const cd* Ta = Sa + SB2;
  const cd* Sc = Sa + ROW_STRIDE;
  const cd* Tc = Sc + SB2;
  for (int e = tid; e < SB2; e += blockDim.x) {
    const int i = e & 31, c = e >> 5;
    W[i * WLD + c] = Ta[e];
    W[(32 + i) * WLD + c] = Sc[e];
    W[i * WLD + 32 + c] = Sa[e];
    W[(32 + i) * WLD + 32 + c] = cd{0.0, 0.0};
    W[i * WLD + 64 + c] = cd{0.0, 0.0};
    W[(32 + i) * WLD + 64 + c] = Tc[e];
  }
  if (tid < 64) prm[tid] = tid;
  __syncthreads();
  // Row equilibration by exact powers of two.  The finite-element rows of different variables
  // scale like dx, 1 and 1/dx; partial pivoting on unscaled rows loses 4-5 digits at 10^4 grid
  // points.  The scales are folded back into the stored factors below.
  if (tid < 64) {
    double mx = 0.0;
    for (int c = 0; c < 96; ++c) mx = fmax(mx, fmax(fabs(W[tid * WLD + c].x), fabs(W[tid * WLD + c].y)));
    rscale[tid] = (mx > 0.0 && isfinite(mx)) ? exp2(-static_cast<double>(ilogb(mx))) : 1.0;
  }
  __syncthreads();
  for (int e = tid; e < 64 * 96; e += blockDim.x) {
    const int i = e / 96, c = e - i * 96;
    W[i * WLD + c] = W[i * WLD + c] * rscale[i];
  }
  __syncthreads();
  // 32 elimination steps, NBF at a time (blocked right-looking LU with look-ahead).  Warp 0 does the
  // sequential part of a block on its own - for each of its columns: bring the column up to date
  // under the earlier pivots of the block, pivot search, row swap, bring the new pivot row up to
  // date, multipliers (warp-synchronous, no CTA barrier) - then all 8 warps apply the NBF rank-1
  // updates in ONE pass over the panel: thread = (column c mod 32, row group).  Same pivots and
  // the same FMA sequence per entry as one step at a time (bit-identical factors), 1 / NBF of the
  // shared-memory traffic and CTA barriers: the updates are bound by shared-memory wavefronts,
  // not by the FP64 pipe (profiles/fp64_pipe_r1.md).  Measured: NBF = 1 / 2 / 4 / 8 -> 2.90 /
  // 2.49 / 2.39 / 2.47 ms per factorisation at G = 10 001 (beyond 4 the serial part dominates).
  const int lane = tid & 31, rg = tid >> 5;
  // partial pivoting over rows k .. 63 of column k (two candidates per lane), row swap; warp 0 only
  auto pivot_and_swap = [&](int k, int nprev) {
    const int i0 = lane, i1 = lane + 32;
    double b0 = i0 >= k ? abs2(W[i0 * WLD + k]) : -1.0;
    const double b1 = abs2(W[i1 * WLD + k]);
    int bi = i0;
    if (b1 > b0) { b0 = b1; bi = i1; }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, b0, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ob > b0 || (ob == b0 && oi < bi)) { b0 = ob; bi = oi; }
    }
    const int pr = bi;   // the butterfly leaves the same (value, index) in every lane
    if (lane == 0 && !(b0 > 0.0)) {   // exactly singular panel column: report like zgbtrf info > 0
      const long long node = (static_cast<long long>(2 * p + 1) << a.level);
      atomicCAS(a.info, 0, static_cast<int>(min(2 * node + 1, static_cast<long long>(a.n))));
      W[pr * WLD + k] = cd{2.2250738585072014e-308, 0.0};
    }
    __syncwarp();
    if (pr != k) {
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const int c = lane + 32 * cc;
        const cd t0 = W[k * WLD + c];
        W[k * WLD + c] = W[pr * WLD + c];
        W[pr * WLD + c] = t0;
      }
      if (lane == 0) {
        const int t0 = prm[k]; prm[k] = prm[pr]; prm[pr] = t0;
        for (int u = 0; u < nprev; ++u) { const cd t1 = lcol[u * 64 + k]; lcol[u * 64 + k] = lcol[u * 64 + pr]; lcol[u * 64 + pr] = t1; }
      }
    }
    __syncwarp();
  };
  // multipliers of column k (rows > k), kept in place and in lc
  auto multipliers = [&](int k, cd* lc) {
    const cd pinv = crecip(W[k * WLD + k]);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = lane + 32 * h;
      if (i > k) {
        const cd l = W[i * WLD + k] * pinv;
        lc[i] = l;
        W[i * WLD + k] = l;
      }
    }
    __syncwarp();
  };
  // ---- look-ahead (default).  The serial part of a block keeps warp 0 busy for about as long as the rank-NBF update
  // keeps all warps busy (ncu: 41 % of the kernel's warp samples were the other seven warps waiting at the barrier in
  // front of the update), so the two now run side by side: while warps 1-7 apply block b to the columns right of the
  // NEXT panel, warp 0 applies it to the next panel's NBF columns and factorises that panel.  Row swaps of a panel are
  // applied inside the panel at once and everywhere else one iteration later, by whoever owns the column, in the same
  // order; the pivot rows of a block are brought up to date column by column just before the rank-NBF update.  Every
  // entry still sees the same operations in the same order: same pivots, bit-identical factors
  // (LGPU_MERGE_LOOKAHEAD=0 runs the loop below; tests/test_gpu_paths.py compares).
  if (LOOKAHEAD) {
    // warp 0: factorise the panel of columns k .. k + NBF - 1 (rows >= k); swaps stay inside the panel
    auto panel_factor = [&](int k, cd* lc, int* pv) {
#pragma unroll
      for (int j = 0; j < NBF; ++j) {
        const int kj = k + j;
        if (j > 0) {   // column k + j under the pivots k .. k + j - 1 of this block
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int i = lane + 32 * h;
            if (i >= kj) {
              cd x = W[i * WLD + kj];
#pragma unroll
              for (int u = 0; u < j; ++u) cfms(x, lc[u * 64 + i], W[(k + u) * WLD + kj]);
              W[i * WLD + kj] = x;
            }
          }
          __syncwarp();
        }
        // pivot search (as pivot_and_swap)
        const int i0 = lane, i1 = lane + 32;
        double b0 = i0 >= kj ? abs2(W[i0 * WLD + kj]) : -1.0;
        const double b1 = abs2(W[i1 * WLD + kj]);
        int bi = i0;
        if (b1 > b0) { b0 = b1; bi = i1; }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, b0, off);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
          if (ob > b0 || (ob == b0 && oi < bi)) { b0 = ob; bi = oi; }
        }
        const int pr = bi;
        if (lane == 0 && !(b0 > 0.0)) {
          const long long node = (static_cast<long long>(2 * p + 1) << a.level);
          atomicCAS(a.info, 0, static_cast<int>(min(2 * node + 1, static_cast<long long>(a.n))));
          W[pr * WLD + kj] = cd{2.2250738585072014e-308, 0.0};
        }
        __syncwarp();
        if (lane == 0) pv[j] = pr;
        if (pr != kj) {
          if (lane < NBF) {   // inside the panel only
            const int c = k + lane;
            const cd t0 = W[kj * WLD + c];
            W[kj * WLD + c] = W[pr * WLD + c];
            W[pr * WLD + c] = t0;
          }
          if (lane == 0) {
            const int t0 = prm[kj]; prm[kj] = prm[pr]; prm[pr] = t0;
            for (int u = 0; u < j; ++u) { const cd t1 = lc[u * 64 + kj]; lc[u * 64 + kj] = lc[u * 64 + pr]; lc[u * 64 + pr] = t1; }
          }
        }
        __syncwarp();
        if (j > 0 && lane < NBF && k + lane > kj) {   // row k + j under the same pivots, panel columns
          const int c = k + lane;
          cd x = W[kj * WLD + c];
#pragma unroll
          for (int u = 0; u < j; ++u) cfms(x, lc[u * 64 + kj], W[(k + u) * WLD + c]);
          W[kj * WLD + c] = x;
        }
        __syncwarp();
        multipliers(kj, lc + j * 64);
      }
    };
    // one thread, one column outside the panel of block k: its row swaps, then its pivot rows under the block's pivots
    auto column_catch_up = [&](int k, const cd* lc, const int* pv, int c, bool update) {
#pragma unroll
      for (int j = 0; j < NBF; ++j) {
        const int pr = pv[j];
        if (pr != k + j) {
          const cd t0 = W[(k + j) * WLD + c];
          W[(k + j) * WLD + c] = W[pr * WLD + c];
          W[pr * WLD + c] = t0;
        }
      }
      if (update) {
#pragma unroll
        for (int j = 1; j < NBF; ++j) {
          cd x = W[(k + j) * WLD + c];
#pragma unroll
          for (int u = 0; u < j; ++u) cfms(x, lc[u * 64 + k + j], W[(k + u) * WLD + c]);
          W[(k + j) * WLD + c] = x;
        }
      }
    };
    // rank-NBF update of rows > kl of one column for the rows i = r0, r0 + rstep, ...
    auto column_update = [&](int k, const cd* lc, int c, int r0, int rstep) {
      const int kl = k + NBF - 1;
      cd wk[NBF];
#pragma unroll
      for (int u = 0; u < NBF; ++u) wk[u] = W[(k + u) * WLD + c];
      for (int i = r0; i < 64; i += rstep) {
        if (i > kl) {
          cd x = W[i * WLD + c];
#pragma unroll
          for (int u = 0; u < NBF; ++u) cfms(x, lc[u * 64 + i], wk[u]);
          W[i * WLD + c] = x;
        }
      }
    };
    if (tid < 32) panel_factor(0, lcol, pivs[0]);
    __syncthreads();
    for (int b = 0; b < SB / NBF; ++b) {
      const int k = b * NBF, kl = k + NBF - 1;
      const cd* lc = lcol + (b & 1) * NBF * 64;
      const int* pv = pivs[b & 1];
      const bool last = b == SB / NBF - 1;
      if (!last && tid < 32) {
        // warp 0: block b on the next panel's columns, then that panel
        if (lane < NBF) column_catch_up(k, lc, pv, kl + 1 + lane, true);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < NBF; ++q) column_update(k, lc, kl + 1 + q, lane, 32);
        __syncwarp();
        panel_factor(k + NBF, lcol + ((b + 1) & 1) * NBF * 64, pivs[(b + 1) & 1]);
      } else {
        // the other warps (all of them for the last block): columns left of the block get its swaps, columns right
        // of the next panel the swaps, the pivot-row update and the rank-NBF update
        const int team0 = last ? 0 : 32, nteam = 256 - team0, tt = tid - team0;
        const int cfirst = last ? kl + 1 : kl + 1 + NBF;           // first column right of the work of warp 0
        for (int c = tt; c < 96; c += nteam) {
          if (c < k) column_catch_up(k, lc, pv, c, false);
          else if (c >= cfirst) column_catch_up(k, lc, pv, c, true);
        }
        if (last) __syncthreads(); else asm volatile("bar.sync 2, 224;" ::: "memory");
        const int nrg = nteam >> 5, myrg = tt >> 5;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          const int c = lane + 32 * cc;
          if (c >= cfirst) column_update(k, lc, c, myrg, nrg);
        }
      }
      __syncthreads();
    }
  } else
  for (int k = 0; k < SB; k += NBF) {
    if (tid < 32) {
#pragma unroll
      for (int j = 0; j < NBF; ++j) {
        const int kj = k + j;
        // column k + j under the pivots k .. k + j - 1 of this block
        if (j > 0) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int i = lane + 32 * h;
            if (i >= kj) {
              cd x = W[i * WLD + kj];
#pragma unroll
              for (int u = 0; u < j; ++u) cfms(x, lcol[u * 64 + i], W[(k + u) * WLD + kj]);
              W[i * WLD + kj] = x;
            }
          }
          __syncwarp();
        }
        pivot_and_swap(kj, j);   // the multipliers of the earlier columns of the block move with their rows
        // row k + j under the same pivots
        if (j > 0) {
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) {
            const int c = lane + 32 * cc;
            if (c > kj) {
              cd x = W[kj * WLD + c];
#pragma unroll
              for (int u = 0; u < j; ++u) cfms(x, lcol[u * 64 + kj], W[(k + u) * WLD + c]);
              W[kj * WLD + c] = x;
            }
          }
          __syncwarp();
        }
        multipliers(kj, lcol + j * 64);
      }
    }
    __syncthreads();
    const int kl = k + NBF - 1;
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      const int c = lane + 32 * cc;
      if (c > kl) {
        cd wk[NBF];
#pragma unroll
        for (int u = 0; u < NBF; ++u) wk[u] = W[(k + u) * WLD + c];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int i = rg + 8 * r;
          if (i > kl) {
            cd x = W[i * WLD + c];
#pragma unroll
            for (int u = 0; u < NBF; ++u) cfms(x, lcol[u * 64 + i], wk[u]);
            W[i * WLD + c] = x;
          }
        }
      }
    }
    __syncthreads();
  }
  // the reduced rows leave first ...
  {
    cd* Sn = a.dst + static_cast<size_t>(p) * ROW_STRIDE;
    cd* Tn = Sn + SB2;
    for (int e = tid; e < SB2; e += blockDim.x) {
      const int i = e & 31, c = e >> 5;
      const double unscale = 1.0 / rscale[prm[32 + i]];
      Sn[e] = W[(32 + i) * WLD + 32 + c] * unscale;
      Tn[e] = W[(32 + i) * WLD + 64 + c] * unscale;
    }
  }
  __syncthreads();
  // ... and their place takes X = L11^-1 (unit lower triangular).  Column sweep: warp w owns columns 4w .. 4w + 3,
  // lane = row; once x_j is final it is broadcast and every row below subtracts L(i, j) x_j - the same subtractions
  // in the same order as forward substitution row by row (bit-identical), but 32 dependent steps per column instead
  // of 496 (one thread per column before: 13 us of the ~83 us a merge takes).
  {
    const int c0 = 4 * rg;
    cd v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = cd{lane == c0 + q ? 1.0 : 0.0, 0.0};
    for (int j = c0; j < SB - 1; ++j) {
      const cd lij = lane > j ? W[lane * WLD + j] : cd{0.0, 0.0};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const cd xj = cd{__shfl_sync(0xffffffffu, v[q].x, j), __shfl_sync(0xffffffffu, v[q].y, j)};
        if (lane > j && j >= c0 + q) cfms(v[q], lij, xj);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) X[lane * WLD + c0 + q] = lane >= c0 + q ? v[q] : cd{0.0, 0.0};
  }
  __syncthreads();
  // M2 = L21 L11^-1 over L21 (rows 32..63, columns 0..31 of W): thread (i, 4 columns)
  {
    const int i = tid & 31, cg = tid >> 5;
    cd m2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = cg * 4 + q;
      cd v{0.0, 0.0};
      for (int k = c; k < SB; ++k) cfma(v, W[(32 + i) * WLD + k], X[k * WLD + c]);
      m2[q] = v;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) W[(32 + i) * WLD + cg * 4 + q] = m2[q];
  }
  __syncthreads();
  cd* rec = a.pairs + static_cast<size_t>(p) * PAIR_STRIDE;
  if (tid < 64) reinterpret_cast<uint8_t*>(rec + PR_PERM)[tid] = static_cast<uint8_t>(prm[tid]);
  for (int e = tid; e < SB2; e += blockDim.x) {
    const int i = e & 31, c = e >> 5;
    // scale of pivot row c folded into column c of L11^-1; reduced rows return to their own scale
    if (i >= c) rec[PR_L11I + tri_lo_off(c) + i - c] = X[i * WLD + c] * rscale[prm[c]];
    // M2 acts on the scaled pivot rows (column c) and returns reduced rows to their own scale
    rec[PR_L21 + e] = W[(32 + i) * WLD + c] * (rscale[prm[c]] / rscale[prm[32 + i]]);
    rec[PR_E + e] = W[i * WLD + 32 + c];
    rec[PR_F + e] = W[i * WLD + 64 + c];
    // U leaves with unit diagonal: row i divided by U_ii, whose reciprocal takes the diagonal slot
    if (i <= c) {
      const cd dinv = crecip(W[i * WLD + i]);
      rec[PR_U + tri_up_off(c) + i] = (i == c) ? dinv : W[i * WLD + c] * dinv;
    }
  }
}

constexpr int TLD = 65;

// Dense top system: rows [node 0 ; last reduced row ; node n_pad - 1] on (z_0, z_K-1).
__global__ void __launch_bounds__(256) slu_top_factor_kernel(FactorArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cd* W = reinterpret_cast<cd*>(smem_raw);       // [64][TLD]
  cd* X = W + 64 * TLD;                          // [64][TLD] inverse of L
  cd* lcol = X + 64 * TLD;
  int* prm = reinterpret_cast<int*>(lcol + 64);
  __shared__ int piv_row;
  __shared__ double rscale[64];
  const int tid = threadIdx.x;
  const int TS = a.top_size;
  const int last = a.n_pad - 1;
  for (int e = tid; e < 64 * 64; e += blockDim.x) {
    const int row = e >> 6, col = e & 63;
    cd v{0.0, 0.0};
    if (row < TS && col < TS) {
      if (row < 16) {                                    // block row 0 on z_0 = (x_0, x_1)
        if (col < 32) v = m_entry(a, 0, col < 16 ? 1 : 2, row, col & 15);
      } else if (TS == 64 && row < 48) {                 // last reduced row: S on z_0, T on z_K-1
        const cd* S = a.src;
        v = col < 32 ? S[col * SB + row - 16] : S[SB2 + (col - 32) * SB + row - 16];
      } else {                                           // block row n_pad-1 on z_K-1
        const int i = row - (TS - 16), c0 = col - (TS - 32);
        if (c0 >= 0) v = m_entry(a, last, c0 < 16 ? 0 : 1, i, c0 & 15);
      }
    } else if (row == col) {
      v = cd{1.0, 0.0};
    }
    W[row * TLD + col] = v;
  }
  if (tid < 64) prm[tid] = tid;
  __syncthreads();
  if (tid < 64) {   // row equilibration, as in the merge kernel
    double mx = 0.0;
    for (int c = 0; c < 64; ++c) mx = fmax(mx, fmax(fabs(W[tid * TLD + c].x), fabs(W[tid * TLD + c].y)));
    rscale[tid] = (mx > 0.0 && isfinite(mx)) ? exp2(-static_cast<double>(ilogb(mx))) : 1.0;
  }
  __syncthreads();
  for (int e = tid; e < 64 * 64; e += blockDim.x) W[(e >> 6) * TLD + (e & 63)] = W[(e >> 6) * TLD + (e & 63)] * rscale[e >> 6];
  __syncthreads();
  for (int k = 0; k < TS; ++k) {
    if (tid < 32) {
      const int i0 = tid, i1 = tid + 32;
      double b0 = (i0 >= k && i0 < TS) ? abs2(W[i0 * TLD + k]) : -1.0;
      const double b1 = (i1 >= k && i1 < TS) ? abs2(W[i1 * TLD + k]) : -1.0;
      int bi = i0;
      if (b1 > b0) { b0 = b1; bi = i1; }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, b0, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ob > b0 || (ob == b0 && oi < bi)) { b0 = ob; bi = oi; }
      }
      if (tid == 0) {
        if (!(b0 > 0.0)) {
          bi = k;
          atomicCAS(a.info, 0, a.n);
          W[k * TLD + k] = cd{2.2250738585072014e-308, 0.0};
        }
        piv_row = bi;
      }
    }
    __syncthreads();
    const int pr = piv_row;
    if (pr != k) {
      if (tid < 64) {
        const cd t0 = W[k * TLD + tid];
        W[k * TLD + tid] = W[pr * TLD + tid];
        W[pr * TLD + tid] = t0;
      } else if (tid == 64) {
        const int t0 = prm[k]; prm[k] = prm[pr]; prm[pr] = t0;
      }
    }
    __syncthreads();
    if (tid > k && tid < TS) {
      const cd l = W[tid * TLD + k] * crecip(W[k * TLD + k]);
      lcol[tid] = l;
      W[tid * TLD + k] = l;
    }
    __syncthreads();
    for (int e = tid; e < 64 * 64; e += blockDim.x) {
      const int i = e >> 6, c = e & 63;
      if (i > k && i < TS && c > k && c < TS) cfms(W[i * TLD + c], lcol[i], W[k * TLD + c]);
    }
    __syncthreads();
  }
  // X = L^-1 by a warp-parallel column sweep (see slu_merge_kernel): warp w owns columns 8w .. 8w + 7, a lane holds
  // rows lane and lane + 32; same subtractions in the same order as row-by-row forward substitution.  One thread per
  // column took ~85 us of this kernel's 185 us (2016 dependent complex FMAs with a shared-memory load each).
  {
    const int lane = tid & 31, c0 = 8 * (tid >> 5);
    cd v0[8], v1[8];   // rows lane, lane + 32
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      v0[q] = cd{lane == c0 + q ? 1.0 : 0.0, 0.0};
      v1[q] = cd{lane + 32 == c0 + q ? 1.0 : 0.0, 0.0};
    }
    for (int j = c0; j < TS - 1; ++j) {
      const cd l0 = lane > j ? W[lane * TLD + j] : cd{0.0, 0.0};
      const cd l1 = (lane + 32 > j && lane + 32 < TS) ? W[(lane + 32) * TLD + j] : cd{0.0, 0.0};
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const cd src = j < 32 ? v0[q] : v1[q];
        const cd xj = cd{__shfl_sync(0xffffffffu, src.x, j & 31), __shfl_sync(0xffffffffu, src.y, j & 31)};
        if (j >= c0 + q) {
          if (lane > j) cfms(v0[q], l0, xj);
          if (lane + 32 > j && lane + 32 < TS) cfms(v1[q], l1, xj);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = c0 + q;
      X[lane * TLD + c] = lane >= c ? v0[q] : cd{0.0, 0.0};
      X[(lane + 32) * TLD + c] = lane + 32 >= c ? v1[q] : cd{0.0, 0.0};
    }
  }
  __syncthreads();
  if (tid < 64) reinterpret_cast<uint8_t*>(a.top)[tid] = static_cast<uint8_t>(prm[tid]);
  for (int e = tid; e < 64 * 64; e += blockDim.x) {
    const int i = e & 63, c = e >> 6;   // column-major, leading dimension 64
    a.top[TOP_LINV + e] = X[i * TLD + c] * rscale[prm[c]];
    cd u{0.0, 0.0};
    if (i < c && c < TS) u = W[i * TLD + c] * crecip(W[i * TLD + i]);
    else if (i == c) u = crecip(W[i * TLD + c]);
    a.top[TOP_U + e] = u;
  }
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

StageArgs make_stage_args(const SluPlan& plan, const SluDevice& d, int s, const cd* b, cd* xv, const RhsEll* ell) {
  const SluStage& st = plan.stages[s];
  StageArgs a{};
  a.l0 = st.l0; a.mu = st.mu; a.m0 = st.m0; a.nchunks = st.nchunks;
  a.n = plan.n; a.n_pad = plan.n_pad; a.K = plan.K; a.top_size = plan.top_size;
  for (int lam = 0; lam < st.mu; ++lam) {
    const SluLevel& lv = plan.levels[st.l0 + lam];
    a.lv[lam] = LevelRef{lv.m, lv.off_pairs};
  }
  a.pairs = d.pairs;
  a.top = d.top;
  a.b = b;
  a.fin = s == 0 ? b + BLK : d.rhs + st.off_fin * SB;   // level-0 row k = b[(2k+1)*16 ...]
  a.fin_is_rhs = s == 0;
  if (ell != nullptr) a.ell = *ell;
  const bool top = s == static_cast<int>(plan.stages.size()) - 1;
  a.fout = top ? nullptr : d.rhs + plan.stages[s + 1].off_fin * SB;
  a.gvec = d.gvec;
  a.xv = xv;
  return a;
}

// ---- shared-memory budgets of the solve kernels
// Three CTAs per SM (232448 B of shared memory, 1 KB reserved per CTA) keep every chunk of the
// first stage resident at once; the narrow upper stages get the whole SM.
constexpr size_t SMEM_PER_CTA_3 = (232448 - 3 * 1024) / 3;
constexpr size_t SMEM_PER_CTA_1 = 232448 - 1024;

struct RingShape {
  int ns, nu;
  size_t bytes;
};
size_t fwd_smem(int mu, int ns) {
  return sizeof(cd) * (static_cast<size_t>(ns) * GRAN + (2 << mu) * SB + 2 * NCW * SB) + 16 * ns;
}
size_t bwd_smem(int mu, int ns, int nu) {
  return sizeof(cd) * (static_cast<size_t>(ns) * GRAN + static_cast<size_t>(nu) * TRI +
                       ((1 << mu) + 1) * SB + 2 * NCW * SB) + 16 * (ns + nu);
}
size_t top_smem(int mu, int ns, int nu) {
  return sizeof(cd) * (static_cast<size_t>(ns) * GRAN + static_cast<size_t>(nu) * TRI + (2 << mu) * SB +
                       ((1 << mu) + 1) * SB + 2 * NCW * SB + 64 * 64 + 64) + 16 * (ns + nu);
}
RingShape fwd_shape(const StageArgs& a) {
  const size_t budget = a.nchunks > 148 ? SMEM_PER_CTA_3 : SMEM_PER_CTA_1;
  int ns = 2;
  while (ns < 8 && fwd_smem(a.mu, ns + 1) <= budget) ++ns;
  return RingShape{ns, 0, fwd_smem(a.mu, ns)};
}
// U slots first (parallel solves per level, ideally twice the widest level so that the next
// batch loads during the current solves), then block slots with what is left
RingShape bwd_shape(const StageArgs& a, bool top) {
  const size_t budget = (!top && a.nchunks > 148) ? SMEM_PER_CTA_3 : SMEM_PER_CTA_1;
  auto bytes = [&](int ns, int nu) { return top ? top_smem(a.mu, ns, nu) : bwd_smem(a.mu, ns, nu); };
  const int widest = std::max(1, (1 << a.mu) / 2);
  const int nu_max = std::min(16, 2 * std::min(widest, NCW));
  int ns = 2, nu = 1;
  while (nu < nu_max && bytes(ns, nu + 1) <= budget) ++nu;
  while (ns < 6 && bytes(ns + 1, nu) <= budget) ++ns;
  return RingShape{ns, nu, bytes(ns, nu)};
}

constexpr size_t MERGE_SMEM = sizeof(cd) * (64 * WLD + 2 * NBF * 64) + sizeof(int) * 64;   // 105 KiB: two CTAs per SM
constexpr size_t TOPF_SMEM = sizeof(cd) * (2 * 64 * TLD + 64) + sizeof(int) * 64;

void configure_kernels() {
  static PerDeviceOnce once;
  once.run([] {
  const int big = static_cast<int>(SMEM_PER_CTA_1);
  CUDA_CHECK(cudaFuncSetAttribute(slu_fwd_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CUDA_CHECK(cudaFuncSetAttribute(slu_bwd_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CUDA_CHECK(cudaFuncSetAttribute(slu_top_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CUDA_CHECK(cudaFuncSetAttribute(slu_fused_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CUDA_CHECK(cudaFuncSetAttribute(slu_upper_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CUDA_CHECK(cudaFuncSetAttribute(slu_merge_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CUDA_CHECK(cudaFuncSetAttribute(slu_merge_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CUDA_CHECK(cudaFuncSetAttribute(slu_top_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  });
}

}  // namespace

void slu_factorize(const SluPlan& plan, const SluDevice& d, cd sigma, cudaStream_t stream,
                   LaunchLog* log) {
  configure_kernels();
  CUDA_CHECK(cudaMemsetAsync(d.info, 0, sizeof(int32_t), stream));
  const int nl = static_cast<int>(plan.levels.size());
  const int m0 = plan.K - 1;
  // read A and B (2 x 12 288 G), write LAPACK-equivalent factors (24 064 G)
  log->begin(LK_FACTOR, 48640.0 * plan.n);
  FactorArgs a{};
  a.A = d.A; a.B = d.B; a.sigma = sigma; a.n = plan.n; a.n_pad = plan.n_pad; a.K = plan.K;
  a.top = d.top; a.top_size = plan.top_size; a.info = d.info; a.padmask = d.padmask;
  auto rows_of = [&](int l) {
    return d.work + (l < nl ? plan.levels[l].off_rows : plan.off_rows_final) * ROW_STRIDE;
  };
  if (m0 >= 1) {
    a.dst = rows_of(0);
    slu_build_rows_kernel<<<m0, 256, 0, stream>>>(a);
    log->launches += 1;
  }
  for (int l = 0; l < nl; ++l) {
    const SluLevel& lv = plan.levels[l];
    a.src = rows_of(l);
    a.dst = rows_of(l + 1);
    a.pairs = d.pairs + lv.off_pairs * PAIR_STRIDE;
    a.m = lv.m;
    a.npairs = lv.npairs;
    a.level = l;
    static const bool lookahead = [] { const char* e = std::getenv("LGPU_MERGE_LOOKAHEAD"); return !(e && e[0] == '0'); }();
    if (lookahead) slu_merge_kernel<true><<<lv.npairs + (lv.m & 1), 256, MERGE_SMEM, stream>>>(a);
    else slu_merge_kernel<false><<<lv.npairs + (lv.m & 1), 256, MERGE_SMEM, stream>>>(a);
    log->launches += 1;
  }
  a.src = rows_of(nl);
  slu_top_factor_kernel<<<1, 256, TOPF_SMEM, stream>>>(a);
  log->end();
  log->launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

// Algorithmic bytes of one stage launch: the share of LAPACK's band-LU factor an equivalent
// zgbtrs sweep streams for the grid points this stage eliminates (forward: 31 multipliers per
// column = 7936 B per grid point, backward: 63 entries of U per column = 16128 B), plus the
// right-hand-side / solution vectors it must touch.
static double stage_algo_bytes(const SluPlan& plan, int s, double per_point) {
  const SluStage& st = plan.stages[s];
  double pairs = 0.0;
  for (int lam = 0; lam < st.mu; ++lam) pairs += plan.levels[st.l0 + lam].npairs;
  return pairs * 2.0 * per_point + 512.0 * (st.m0 + st.nchunks);
}

// first stage (>= 1) from which the rest of the solve runs as one cooperative launch, or the
// index of the top stage when there is nothing to fuse
static int first_fused_stage(const SluPlan& plan) {
  const int sm_count = device_sm_count();
  static const bool enabled = [] { const char* e = std::getenv("LGPU_SLU_FUSE"); return !(e && e[0] == '0'); }();
  const int ns = static_cast<int>(plan.stages.size());
  if (!enabled || ns < 3) return ns - 1;
  int sf = std::max(1, ns - MAX_FUSED);
  while (sf < ns - 1 && plan.stages[sf].nchunks > sm_count) ++sf;   // one CTA per SM, all resident
  return sf;
}

// Same for the resident-record kernel (slu_upper_kernel): every chunk of every fused stage gets its own
// CTA, so the CTAs of ALL fused stages must be co-resident; chunks of at most three pairs (mu <= 2), a
// top stage of at most two rows.  -1: not applicable to this plan.
static int first_upper_stage(const SluPlan& plan) {
  static const bool enabled = [] { const char* e = std::getenv("LGPU_SLU_UPPER"); return !(e && e[0] == '0'); }();
  const int ns = static_cast<int>(plan.stages.size());
  if (!enabled || ns < 2 || plan.top_size != 64) return -1;
  const SluStage& top = plan.stages[ns - 1];
  if (top.m0 > 2 || top.mu > 1) return -1;
  const int sm_count = device_sm_count();
  int sf = ns - 1, ctas = 1;
  while (sf > 1 && ns - (sf - 1) <= MAX_FUSED && plan.stages[sf - 1].mu <= 2 &&
         ctas + plan.stages[sf - 1].nchunks <= sm_count) {
    --sf;
    ctas += plan.stages[sf].nchunks;
  }
  return sf;
}

// SMs a solve with this plan may hold while it WAITS: the CTAs of the upper-stage launch poll mailboxes written by
// other CTAs of the same launch (all of them must be resident), and the programmatically launched first-stage kernel
// behind it sits on its SMs until that launch completes.  One context (one stream) cannot deadlock - the launches
// are ordered - but several contexts of one process could each get a part of their upper-stage CTAs resident and
// wait for the rest for ever; the C ABI therefore admits concurrent solves only while the sum of their demands
// fits the device (api.cu: SolveTicket).
int slu_coresident_demand(const SluPlan& plan) {
  const int ns = static_cast<int>(plan.stages.size());
  if (ns < 2) return 1;
  const int su = first_upper_stage(plan);
  int ctas = 0;
  if (su >= 0) {
    for (int s = su; s < ns; ++s) ctas += s == ns - 1 ? 1 : plan.stages[s].nchunks;
  } else {
    const int sf = first_fused_stage(plan);
    ctas = sf < ns - 1 ? plan.stages[sf].nchunks : 1;   // cooperative launch: the driver guarantees residency itself
  }
  return ctas + (plan.stages[0].nchunks + 2) / 3;
}

// Programmatic dependent launch of the solve's kernels: a kernel may start while its predecessor on
// the stream is still running; it copies its first factor records (static data) and then blocks in
// griddepcontrol.wait until the predecessor has completed.
// edge: 0 forward stage kernels, 1 upper-stage kernel, 2 backward stage kernels (LGPU_PDL_MASK: debugging, default all)
static bool pdl_enabled(int edge) {
  static const bool on = [] { const char* e = std::getenv("LGPU_PDL"); return !(e && e[0] == '0'); }();
  static const int mask = [] { const char* e = std::getenv("LGPU_PDL_MASK"); return e ? std::atoi(e) : 7; }();
  return on && ((mask >> edge) & 1);
}
template <typename... Args>
static void launch_pdl(int edge, void (*kernel)(Args...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled(edge) ? 1 : 0;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, args...));
}

void slu_solve(const SluPlan& plan, const SluDevice& d, const cd* b, cd* x, cudaStream_t stream,
               LaunchLog* log, const RhsEll* ell, bool x_padded) {
  configure_kernels();
  const int ns = static_cast<int>(plan.stages.size());
  const int su = first_upper_stage(plan);
  const int sf = su >= 0 ? su : first_fused_stage(plan);
  cd* xv = (plan.n_pad == plan.n || x_padded) ? x : d.xpad;
  for (int s = 0; s < sf; ++s) {
    const StageArgs a = make_stage_args(plan, d, s, b, xv, ell);
    // with the B v product fused in: plus the band of B and the vector v (12 288 + 256 bytes per grid point)
    log->begin(s == 0 ? LK_FWD0 : LK_FWD,
               stage_algo_bytes(plan, s, 7936.0) + (s == 0 && ell != nullptr ? 12544.0 * plan.n : 0.0));
    const RingShape sh = fwd_shape(a);
    launch_pdl(0, slu_fwd_stage_kernel, dim3(a.nchunks), dim3(RING_THREADS), sh.bytes, stream, a, sh.ns);
    log->end();
  }
  double top_bytes = stage_algo_bytes(plan, ns - 1, 24064.0) + 2.0 * 24064.0 * 2;
  if (su < 0 && sf == ns - 1) {
    const StageArgs a = make_stage_args(plan, d, ns - 1, b, xv, ell);
    log->begin(LK_TOP, top_bytes);
    const RingShape sh = bwd_shape(a, true);
    slu_top_stage_kernel<<<1, RING_THREADS, sh.bytes, stream>>>(a, sh.ns, sh.nu);
    log->end();
  } else {
    // stage arguments shared by the two fused kernels: mailboxes between the CTAs of the launch
    StageArgs fst[MAX_FUSED];
    int mu_max = 0;
    const unsigned long long parity = (*d.epoch + 1) & 1ull;   // of the solve being launched
    const size_t half = slu_mbox_half(plan);
    cd* box = d.mbox + parity * half;
    cd* box_other = d.mbox + (1 - parity) * half;
    const size_t zoff = plan.rhs_vecs * SB;
    for (int s = sf; s < ns; ++s) {
      StageArgs& a = fst[s - sf];
      a = make_stage_args(plan, d, s, b, xv, ell);
      if (s > sf) {            // input rows come from another CTA of this launch
        a.fin = box + plan.stages[s].off_fin * SB;
        a.poll_in = 1;
      }
      if (s < ns - 1) {        // output rows go to another CTA of this launch
        a.fout = box + plan.stages[s + 1].off_fin * SB;
        a.fout_clear = box_other + plan.stages[s + 1].off_fin * SB;
        a.poll_z = 1;          // ... and its boundary unknowns come from one
      }
      a.zbox = box + zoff;
      a.zbox_clear = box_other + zoff;
      mu_max = std::max(mu_max, plan.stages[s].mu);
      if (s < ns - 1) top_bytes += stage_algo_bytes(plan, s, 24064.0);
    }
    ++*d.epoch;
    log->begin(LK_TOP, top_bytes);
    if (su >= 0) {
      UpperArgs u{};
      u.nst = ns - sf;
      int ctas = 0;
      for (int i = 0; i < u.nst; ++i) {
        u.st[i] = fst[i];
        u.cta0[i] = ctas;
        ctas += i == u.nst - 1 ? 1 : fst[i].nchunks;
      }
      u.cta0[u.nst] = ctas;
      // L2 prefetch for the backward first-stage kernel: its upper levels, as many as fit the budget (the lowest
      // level is half of all records and is left to stream from HBM while the prefetched levels are consumed).
      // OFF by default (LGPU_SLU_L2PF_MB = 0): measured at G = 10 001 with 90 MB prefetched, the backward first-stage
      // kernel gained 2.6 us but the mailbox round trips of this launch, queued behind the prefetch traffic, lost 6 us.
      static const double pf_mb = [] { const char* e = std::getenv("LGPU_SLU_L2PF_MB"); return e ? atof(e) : 0.0; }();
      if (sf == 1 && u.nst >= 2) {
        const SluStage& s0 = plan.stages[0];
        int lo = s0.mu;   // prefetch levels [lo, mu)
        size_t recs = 0;
        while (lo > 0 && (recs + plan.levels[lo - 1].npairs) * sizeof(cd) * UP_BWD_ELEMS <= pf_mb * 1.0e6) {
          --lo;
          recs += plan.levels[lo].npairs;
        }
        if (recs > 0) {
          u.pf_first = d.pairs + plan.levels[lo].off_pairs * PAIR_STRIDE;
          u.pf_count = static_cast<int>(recs);
        }
      }
      launch_pdl(1, slu_upper_kernel, dim3(ctas), dim3(UP_THREADS), UP_SMEM, stream, u);
    } else {
      FusedArgs f{};
      f.nst = ns - sf;
      for (int i = 0; i < f.nst; ++i) f.st[i] = fst[i];
      f.cmax = 1 << mu_max;
      StageArgs widest = f.st[0];
      widest.mu = mu_max;
      const RingShape sh = bwd_shape(widest, true);
      f.ns = sh.ns; f.nu = sh.nu;
      void* args[] = {&f};
      CUDA_CHECK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(slu_fused_stage_kernel), dim3(f.st[0].nchunks),
                                             dim3(RING_THREADS), args, sh.bytes, stream));
    }
    log->end();
  }
  if (d.signal != nullptr) {
    ++d.signal->epoch;
    d.signal->last = 0;
  }
  for (int s = sf - 1; s >= 0; --s) {
    StageArgs a = make_stage_args(plan, d, s, b, xv, ell);
    if (s == 0 && d.signal != nullptr && d.signal->flags != nullptr && xv == x && a.nchunks <= d.signal->nchunks) {
      a.done_flags = d.signal->flags;
      a.done_epoch = d.signal->epoch;
      d.signal->last = d.signal->epoch;
      d.signal->mu = a.mu;
      d.signal->x = x;
    }
    log->begin(s == 0 ? LK_BWD0 : LK_BWD, stage_algo_bytes(plan, s, 16128.0));
    const RingShape sh = bwd_shape(a, false);
    launch_pdl(2, slu_bwd_stage_kernel, dim3(a.nchunks), dim3(RING_THREADS), sh.bytes, stream, a, sh.ns, sh.nu);
    log->end();
  }
  log->launches += 2 * sf + 1;
  if (xv != x) {
    CUDA_CHECK(cudaMemcpyAsync(x, xv, sizeof(cd) * plan.n * BLK, cudaMemcpyDeviceToDevice, stream));
  }
  CUDA_CHECK(cudaGetLastError());
}

#ifdef LGPU_TRACE
extern "C" int lgpu_debug_solve_trace(int which, unsigned long long* out) {
  return static_cast<int>(cudaMemcpyFromSymbol(out, g_trace, sizeof(unsigned long long) * TRACE_CTAS * TRACE_SLOTS,
                                               sizeof(unsigned long long) * TRACE_CTAS * TRACE_SLOTS * which));
}
#endif

static inline int use_a_count(cd v) { return (v.x != 0.0 || v.y != 0.0) ? 1 : 0; }

void block_matvec(int n, const cd* A, const cd* B, cd aa, cd ab, const cd* x, const cd* z, cd* y,
                  cudaStream_t stream, LaunchLog* log) {
  const int ctas = (n + 7) / 8;
  log->begin(LK_MATVEC, ((use_a_count(aa) + use_a_count(ab)) * 12288.0 + (z ? 768.0 : 512.0)) * n);
  const bool use_a = aa.x != 0.0 || aa.y != 0.0, use_b = ab.x != 0.0 || ab.y != 0.0;
  if (use_a && use_b) block_matvec_kernel<true, true><<<ctas, 256, 0, stream>>>(n, A, B, aa, ab, x, z, y);
  else if (use_a) block_matvec_kernel<true, false><<<ctas, 256, 0, stream>>>(n, A, B, aa, ab, x, z, y);
  else block_matvec_kernel<false, true><<<ctas, 256, 0, stream>>>(n, A, B, aa, ab, x, z, y);
  log->end();
  log->launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace lgpu
