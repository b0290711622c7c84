// C ABI of liblegolas_b200 (include/legolas_b200.h): context, device buffers, and the CUDA
// implementation of the KrylovOps used by the implicitly restarted Arnoldi driver.
#include <cuda_runtime.h>

#include <chrono>
#include <atomic>
#include <condition_variable>
#include <thread>
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "../../include/legolas_b200.h"
#include "arnoldi.cuh"
#include "assemble.cuh"
#include "slu.cuh"
#include "bsparse.cuh"
#include "efs.cuh"
#include "common.cuh"
#include "iram.hpp"

using namespace lgpu;

namespace {

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  void ensure(size_t count) {
    if (count <= cap) return;
    release();
    CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    cap = count;
    // debugging aid: LGPU_DBG_POISON=<byte> fills every fresh device allocation with that byte (255: NaNs), so that
    // a read of memory nothing has written yet shows up in the results instead of depending on what the allocator returns
    static const int poison = [] { const char* e = std::getenv("LGPU_DBG_POISON"); return e ? std::atoi(e) : -1; }();
    if (poison >= 0) CUDA_CHECK(cudaMemset(p, poison & 0xff, std::max<size_t>(count, 1) * sizeof(T)));
  }
};

template <typename T>
struct PinnedBuf {
  T* p = nullptr;
  size_t cap = 0;
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
  void ensure(size_t count) {
    if (count <= cap) return;
    if (p) cudaFreeHost(p);
    p = nullptr;
    CUDA_CHECK(cudaMallocHost(&p, std::max<size_t>(count, 1) * sizeof(T)));
    cap = count;
  }
};

int env_int(const char* name, int fallback) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : fallback;
}

// debugging aid: LGPU_DBG_SERIAL=<mask> lets only one context of the process at a time be in a phase
// (bit 0 assembly, bit 1 factorisation, bit 2 Arnoldi iteration) - which phase is sensitive to neighbours on the GPU?
struct DbgSerial {
  std::mutex* m = nullptr;
  explicit DbgSerial(int phase) {
    static const int mask = [] { const char* e = std::getenv("LGPU_DBG_SERIAL"); return e ? std::atoi(e) : 0; }();
    static std::mutex mu[3];
    if ((mask >> phase) & 1) { m = &mu[phase]; m->lock(); }
  }
  ~DbgSerial() { if (m) m->unlock(); }
};

inline bool dbg_serial_steps() {
  static const bool on = [] { const char* e = std::getenv("LGPU_DBG_SERIAL"); return e && (std::atoi(e) & 8); }();
  return on;
}
inline std::mutex& dbg_step_mutex() { static std::mutex m; return m; }
// debugging aid: LGPU_DBG_DUAL=1 runs every operator application and every fused Gram-Schmidt step of the Arnoldi
// driver twice on the same inputs and counts the 8-byte words in which the two outputs differ
__global__ void dbg_compare_kernel(const unsigned long long* a, const unsigned long long* b, size_t n64,
                                   unsigned long long* counter) {
  unsigned long long bad = 0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n64;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    bad += a[i] != b[i];
  if (bad) atomicAdd(counter, bad);
}
inline bool dbg_dual() {
  static const bool on = std::getenv("LGPU_DBG_DUAL") != nullptr;
  return on;
}
inline void dbg_compare(const cd* a, const cd* b, size_t n, unsigned long long* counter, cudaStream_t stream) {
  dbg_compare_kernel<<<64, 256, 0, stream>>>(reinterpret_cast<const unsigned long long*>(a),
                                             reinterpret_cast<const unsigned long long*>(b), 2 * n, counter);
}

double now_ms() {
  using clk = std::chrono::steady_clock;
  return std::chrono::duration<double, std::milli>(clk::now().time_since_epoch()).count();
}

}  // namespace

struct lgpu_ctx {
  int device = 0;
  int log_level = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int demand_G = -1, demand = 0;   // SMs a call of this context may hold while waiting, for G grid points (SolveTicket)
  int sm_limit = 0;                // lgpu_set_sm_limit: cap on the CTAs of the fused Gram-Schmidt step (0: all SMs)
  std::string err;

  // matrices
  lgpu_settings settings{};
  int G = 0;            // block rows (gridpts)
  int N = 0;            // device dimension: 16 rows per grid point for every physics type
  // hd / hd-1d (src/settings/mod_settings.f08:69-86): the reference numbers 2 nb_eqs rows per grid
  // point; on the device an absent variable leaves an empty slot.  cmap[r] = device row (within a
  // block) of compact row r; all vectors and indices crossing the C ABI use the compact numbering.
  int dsub = BLK;
  int cmap[BLK] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};
  uint32_t padmask = 0; // device rows of a block that belong to no variable
  int Nc() const { return G * dsub; }
  bool compact() const { return dsub != BLK; }
  DevBuf<cd> cbuf;      // staging for compact <-> device layout translation
  bool have[2] = {false, false};
  DevBuf<cd> A, B;
  DevBuf<uint32_t> masks, natmasks;
  DevBuf<double> d_grid, d_gauss, d_fields;
  DevBuf<int32_t> d_plan_i32;
  DevBuf<PairItem> d_items;

  // factorisation
  SluPlan splan;
  DevBuf<cd> pairs, topfac, fwork, rhs, gvec, xpad;
  DevBuf<int32_t> d_info;
  DevBuf<double> grid_copy;     // base grid of the last assembly (eigenfunction assembly)
  DevBuf<cd> ef_in, ef_out;
  DevBuf<int32_t> ef_idx;
  DevBuf<cd> bell_val;          // compressed copy of B for the operator application
  DevBuf<int32_t> bell_col, bell_width;
  DevBuf<double> bell_rval;
  bool bell_real = false;       // every entry of B is real: the product streams the 8-byte copy
  int bell_w = -1;              // longest row of B; -1: not built for the current B
  bool have_grid = false;       // grid_copy matches the resident matrices
  DevBuf<cd> mbox;              // mailboxes of the fused solve stages (slu.cuh)
  DevBuf<unsigned long long> solve_flags;   // SolveSignal::flags
  SolveSignal signal;
  unsigned long long solve_epoch = 0;
  bool factorized = false;
  bool factor_of_B = false;     // general mode: the resident factors are those of B, not of A - sigma B
  cd sigma{0.0, 0.0};
  int lu_info = 0;

  // vectors / Krylov storage
  DevBuf<cd> vx, vy, vu, vr, ve;
  DevBuf<cd> V, vcur, resid, Hdev, Qdev, Z, kpartial, khwork;
  BasisLayout basis{};
  DevBuf<double> kscal;
  DevBuf<unsigned int> kticket;
  DevBuf<unsigned long long> kgbar;
  unsigned long long kgbar_count = 0;
  PinnedBuf<cd> h_stage, h_upload;
  DevBuf<cd> dbg_buf;                       // LGPU_DBG_DUAL scratch
  DevBuf<unsigned long long> dbg_cnt;       // LGPU_DBG_DUAL mismatch counters
  PinnedBuf<double> h_scal;

  LaunchLog log;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_block = nullptr;   // blocking-sync event of stream_sync()
  double t_assemble = 0, t_factor = 0, t_iter = 0, t_extract = 0;

  bool assembled() const { return have[0] && have[1]; }
  SluDevice sdev() {
    SluDevice d{};
    d.A = factor_of_B ? B.p : A.p; d.B = B.p; d.pairs = pairs.p; d.top = topfac.p; d.work = fwork.p; d.rhs = rhs.p;
    d.gvec = gvec.p; d.xpad = xpad.p; d.info = d_info.p;
    d.epoch = &solve_epoch; d.padmask = padmask; d.mbox = mbox.p;
    signal.flags = solve_flags.p;
    signal.nchunks = static_cast<int>(solve_flags.cap);
    d.signal = &signal;
    return d;
  }
};

// Wait for the context's stream.  Contexts that share their GPU (lgpu_set_sm_limit > 0: several host threads drive
// several contexts per process) sleep on a blocking event instead of spinning when the host is short of cores -
// waiting threads of all ranks of the node (LOCAL_WORLD_SIZE processes x sharing contexts each) x 2 > hardware threads -
// so that they do not take the cores from the threads that have launches to issue; with cores to spare spinning is
// faster (measured, 256-unit sweep, three contexts, 32 cores: 0.66 s spinning, 0.78 s blocking).
inline std::atomic<int>& sharing_contexts() { static std::atomic<int> n{0}; return n; }
inline bool blocking_waits() {
  if (const char* e = std::getenv("LGPU_BLOCKING_SYNC")) return e[0] != '0';
  int ranks = 1;
  if (const char* e = std::getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, std::atoi(e));
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  return 2u * static_cast<unsigned>(ranks * std::max(1, sharing_contexts().load())) > hw;
}
inline cudaError_t stream_sync(lgpu_ctx* c) {
  if (c->sm_limit <= 0 || !blocking_waits()) return cudaStreamSynchronize(c->stream);
  if (!c->ev_block) {
    const cudaError_t e = cudaEventCreateWithFlags(&c->ev_block, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  const cudaError_t e = cudaEventRecord(c->ev_block, c->stream);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(c->ev_block);
}

namespace {

// Admission of concurrent calls from several contexts of one process on one device (one host thread per context,
// e.g. a wavenumber sweep that keeps several small units in flight per GPU): a call runs while the SMs its solves may
// hold waiting (slu_coresident_demand) fit the device next to those of the calls already running; a call whose
// demand alone exceeds the device runs alone.  Every entry point synchronises its stream before it returns, so the
// ticket covers all the launches of the call.
struct SolveAdmission {
  std::mutex m;
  std::condition_variable cv;
  int in_use[64] = {};
  static SolveAdmission& get() { static SolveAdmission a; return a; }
};
class SolveTicket {
 public:
  SolveTicket(int dev, int demand) : dev_(dev >= 0 && dev < 64 ? dev : 0), demand_(demand) {
    if (demand_ <= 0) return;
    SolveAdmission& a = SolveAdmission::get();
    const int budget = device_sm_count();
    std::unique_lock<std::mutex> lk(a.m);
    a.cv.wait(lk, [&] { return a.in_use[dev_] == 0 || a.in_use[dev_] + demand_ <= budget; });
    a.in_use[dev_] += demand_;
  }
  ~SolveTicket() {
    if (demand_ <= 0) return;
    SolveAdmission& a = SolveAdmission::get();
    { std::lock_guard<std::mutex> lk(a.m); a.in_use[dev_] -= demand_; }
    a.cv.notify_all();
  }
  SolveTicket(const SolveTicket&) = delete;
  SolveTicket& operator=(const SolveTicket&) = delete;
 private:
  int dev_, demand_;
};

int solve_demand(lgpu_ctx* c) {
  if (c->G <= 0) return 0;
  if (c->demand_G != c->G) {
    // the upper solve stages (mailboxes between their CTAs) and the fused Gram-Schmidt step (device-wide barrier)
    // never run at the same time on one stream: the larger of the two
    const int ntiles = (c->G * BLK + KRYLOV_TILE - 1) / KRYLOV_TILE;
    c->demand = std::max(slu_coresident_demand(make_slu_plan(c->G, env_int("LGPU_SLU_MU0", 4), env_int("LGPU_SLU_MU1", 2),
                                                             env_int("LGPU_SLU_TOP", 2))),
                         krylov_cgs2_grid(ntiles, c->sm_limit));
    c->demand_G = c->G;
  }
  return c->demand;
}

template <typename F>
int guarded(lgpu_ctx* ctx, F&& body) {
  if (!ctx) return LGPU_EINVAL;
  try {
    CUDA_CHECK(cudaSetDevice(ctx->device));
    SolveTicket ticket(ctx->device, solve_demand(ctx));
    return body();
  } catch (const CudaError& e) {
    ctx->err = e.what();
    return LGPU_ENOGPU;
  } catch (const std::bad_alloc&) {
    ctx->err = "host allocation failed";
    return LGPU_ENOMEM;
  } catch (const std::exception& e) {
    ctx->err = e.what();
    return LGPU_EINVAL;
  }
}

int fail(lgpu_ctx* ctx, int code, const std::string& msg) {
  ctx->err = msg;
  return code;
}

struct RowMap { int r[BLK]; };

// dst (device layout, ld = 16 G) <- src (compact, ld = dsub G); empty slots are zeroed
__global__ void expand_rows_kernel(const cd* __restrict__ src, cd* __restrict__ dst, int G, int dsub, RowMap inv,
                                   int ncols) {
  const size_t n = static_cast<size_t>(G) * BLK, total = n * ncols;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t col = i / n, row = i % n;
    const int q = inv.r[row % BLK];
    dst[i] = q >= 0 ? src[col * (static_cast<size_t>(G) * dsub) + (row / BLK) * dsub + q] : cd{0.0, 0.0};
  }
}

// dst (compact) <- src (device layout)
__global__ void compact_rows_kernel(const cd* __restrict__ src, cd* __restrict__ dst, int G, int dsub, RowMap fwd,
                                    int ncols) {
  const size_t nc = static_cast<size_t>(G) * dsub, total = nc * ncols;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t col = i / nc, row = i % nc;
    dst[i] = src[col * (static_cast<size_t>(G) * BLK) + (row / dsub) * BLK + fwd.r[row % dsub]];
  }
}

int map_grid(size_t total) { return static_cast<int>(std::min<size_t>((total + 255) / 256, 148 * 8)); }

// ncols vectors in the reference's numbering (host or device memory) -> device layout
void vec_in(lgpu_ctx* c, const void* src, bool src_on_device, cd* dst, int ncols = 1) {
  const cudaMemcpyKind kind = src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  if (!c->compact()) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, sizeof(cd) * c->N * ncols, kind, c->stream));
    return;
  }
  const size_t cnt = static_cast<size_t>(c->Nc()) * ncols;
  const cd* from = static_cast<const cd*>(src);
  if (!src_on_device) {
    c->cbuf.ensure(cnt);
    CUDA_CHECK(cudaMemcpyAsync(c->cbuf.p, src, sizeof(cd) * cnt, kind, c->stream));
    from = c->cbuf.p;
  }
  RowMap inv;
  for (int i = 0; i < BLK; ++i) inv.r[i] = -1;
  for (int q = 0; q < c->dsub; ++q) inv.r[c->cmap[q]] = q;
  const size_t total = static_cast<size_t>(c->N) * ncols;
  expand_rows_kernel<<<map_grid(total), 256, 0, c->stream>>>(from, dst, c->G, c->dsub, inv, ncols);
  CUDA_CHECK(cudaGetLastError());
  c->log.launches += 1;
}

// device layout -> ncols vectors in the reference's numbering (host or device memory); the copy to
// the host is asynchronous on the context's stream like the plain one
void vec_out(lgpu_ctx* c, const cd* src, void* dst, bool dst_on_device, int ncols = 1) {
  const cudaMemcpyKind kind = dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (!c->compact()) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, sizeof(cd) * c->N * ncols, kind, c->stream));
    return;
  }
  const size_t cnt = static_cast<size_t>(c->Nc()) * ncols;
  cd* to = static_cast<cd*>(dst);
  if (!dst_on_device) {
    c->cbuf.ensure(cnt);
    to = c->cbuf.p;
  }
  RowMap fwd;
  for (int i = 0; i < BLK; ++i) fwd.r[i] = i < c->dsub ? c->cmap[i] : 0;
  compact_rows_kernel<<<map_grid(cnt), 256, 0, c->stream>>>(src, to, c->G, c->dsub, fwd, ncols);
  CUDA_CHECK(cudaGetLastError());
  c->log.launches += 1;
  if (!dst_on_device) CUDA_CHECK(cudaMemcpyAsync(dst, to, sizeof(cd) * cnt, kind, c->stream));
}

void ensure_vectors(lgpu_ctx* c) {
  const size_t n = static_cast<size_t>(c->N);
  c->vx.ensure(n); c->vy.ensure(n); c->vu.ensure(n); c->vr.ensure(n); c->ve.ensure(n);
}

void ensure_krylov_work(lgpu_ctx* c) {
  // one partial row per tile of the basis (dot / update kernels run one CTA per tile)
  const size_t tiles = static_cast<size_t>(make_basis_layout(c->N, 1).ntiles);
  c->kpartial.ensure((3 * tiles + 8) * (KRYLOV_MAXCOL + 1));   // three partial sets for the fused step
  c->khwork.ensure(KRYLOV_MAXCOL + 1);
  if (!c->kscal.p) {
    c->kscal.ensure(4);
    c->kticket.ensure(1);
    CUDA_CHECK(cudaMemsetAsync(c->kticket.p, 0, sizeof(unsigned int), c->stream));
    c->kgbar.ensure(1);
    CUDA_CHECK(cudaMemsetAsync(c->kgbar.p, 0, sizeof(unsigned long long), c->stream));
    c->kgbar_count = 0;
  }
  c->h_scal.ensure(4);
}

KrylovWork kwork(lgpu_ctx* c) {
  KrylovWork w{};
  w.partial = c->kpartial.p; w.hwork = c->khwork.p; w.scal = c->kscal.p; w.ticket = c->kticket.p;
  w.gbar = c->kgbar.p; w.gbar_count = &c->kgbar_count;
  w.grid_cap = c->sm_limit;
  return w;
}

int do_assemble(lgpu_ctx* c, const lgpu_settings* s, const double* d_grid, const double* d_gauss,
                const FieldPtrs& fields) {
  if (s->physics_type < 0 || s->physics_type > 2)
    return fail(c, LGPU_EINVAL, "physics_type must be 0 (mhd), 1 (hd) or 2 (hd-1d)");
  DbgSerial dbg_serial(0);
  const int G = s->gridpts;
  c->settings = *s;
  c->G = G;
  c->N = G * BLK;
  {
    int slots[8];
    const int nb_eqs = state_positions(s->physics_type, slots);
    c->dsub = 2 * nb_eqs;
    c->padmask = 0xffffu;
    for (int q = 0; q < nb_eqs; ++q) {
      c->cmap[2 * q] = 2 * slots[q];
      c->cmap[2 * q + 1] = 2 * slots[q] + 1;
      c->padmask &= ~(3u << (2 * slots[q]));
    }
  }
  c->factorized = false;
  const size_t nblk = static_cast<size_t>(G) * 3 * BLK2;
  c->A.ensure(nblk);
  c->B.ensure(nblk);
  c->masks.ensure(static_cast<size_t>(G) * MASK_WORDS);
  c->natmasks.ensure(2 * 2 * 4 * 8 + 2);
  CUDA_CHECK(cudaMemsetAsync(c->natmasks.p, 0, (2 * 2 * 4 * 8 + 2) * sizeof(uint32_t), c->stream));

  AsmParams p{};
  p.gridpts = G;
  p.geometry = s->geometry;
  p.k2 = s->k2; p.k3 = s->k3;
  p.gamma_1 = (s->incompressible ? 1.0e12 : s->gamma) - 1.0;
  p.mu = s->viscosity_value;
  p.efrac = s->electron_fraction;
  static const double nodes[4] = {-0.861136311594053, -0.339981043584856, 0.339981043584856,
                                  0.861136311594053};
  static const double weights[4] = {0.347854845137454, 0.652145154862546, 0.652145154862546,
                                    0.347854845137454};
  bool custom = false;
  for (int i = 0; i < 4; ++i) custom = custom || s->gauss_weights[i] != 0.0;
  for (int i = 0; i < 4; ++i) {
    p.nodes[i] = custom ? s->gauss_nodes[i] : nodes[i];
    p.weights[i] = custom ? s->gauss_weights[i] : weights[i];
  }

  // plans -> device
  const TermPlan pe = build_term_plan(*s, false), pn = build_term_plan(*s, true);
  const std::vector<int32_t> essl = essential_indices(*s, false), essr = essential_indices(*s, true);
  std::vector<int32_t> packed;
  auto push = [&](const std::vector<int32_t>& v) {
    const size_t off = packed.size();
    packed.insert(packed.end(), v.begin(), v.end());
    return off;
  };
  const size_t o_sb_e = push(pe.slot_begin), o_t_e = push(pe.term_ids);
  const size_t o_sb_n = push(pn.slot_begin), o_t_n = push(pn.term_ids);
  const size_t o_el = push(essl), o_er = push(essr);
  c->d_plan_i32.ensure(packed.size());
  CUDA_CHECK(cudaMemcpyAsync(c->d_plan_i32.p, packed.data(), packed.size() * sizeof(int32_t),
                             cudaMemcpyHostToDevice, c->stream));
  std::vector<PairItem> items = pe.items;
  items.insert(items.end(), pn.items.begin(), pn.items.end());
  c->d_items.ensure(items.size());
  CUDA_CHECK(cudaMemcpyAsync(c->d_items.p, items.data(), items.size() * sizeof(PairItem),
                             cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK(stream_sync(c));   // host vectors go out of scope below
  DevicePlan de{pe.nslots(), static_cast<int32_t>(pe.items.size()), c->d_plan_i32.p + o_sb_e,
                c->d_plan_i32.p + o_t_e, c->d_items.p};
  DevicePlan dn{pn.nslots(), static_cast<int32_t>(pn.items.size()), c->d_plan_i32.p + o_sb_n,
                c->d_plan_i32.p + o_t_n, c->d_items.p + pe.items.size()};

  c->log.stream = c->stream;
  CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
  {   // sampled fields + grids in, two block-tridiagonal matrices out
    int nf = 0;
    for (int f = 0; f < NFIELD; ++f) nf += fields.f[f] != nullptr;
    c->log.begin(LK_ASSEMBLE, 8.0 * nf * 4.0 * (G - 1) + 8.0 * (G + 4.0 * (G - 1)) + 2.0 * 12288.0 * G);
  }
  launch_assemble(p, de, fields, d_grid, d_gauss, c->A.p, c->B.p, c->masks.p, c->stream);
  launch_boundaries(p, dn, fields, d_grid, d_gauss, c->A.p, c->B.p, c->masks.p, c->natmasks.p,
                    c->d_plan_i32.p + o_el, static_cast<int>(essl.size()), c->d_plan_i32.p + o_er,
                    static_cast<int>(essr.size()), c->stream);
  c->log.end();
  c->log.launches += 2;
  CUDA_CHECK(cudaEventRecord(c->ev1, c->stream));
  CUDA_CHECK(cudaEventSynchronize(c->ev1));
  float ms = 0.f;
  CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->t_assemble = ms;
  c->have[0] = c->have[1] = true;
  c->bell_w = -1;
  c->grid_copy.ensure(G);
  CUDA_CHECK(cudaMemcpyAsync(c->grid_copy.p, d_grid, sizeof(double) * G, cudaMemcpyDeviceToDevice, c->stream));
  c->have_grid = true;
  return LGPU_OK;
}

// factors of A - sigma B, or (of_B, general mode: sigma ignored) of B itself
int do_factorize(lgpu_ctx* c, cd sigma, bool of_B = false) {
  if (!c->assembled()) return fail(c, LGPU_ESTATE, "factorize: matrices not assembled");
  DbgSerial dbg_serial(1);
  c->factor_of_B = of_B;
  if (of_B) sigma = cd{0.0, 0.0};
  if (c->splan.n != c->G) {
    c->splan = make_slu_plan(c->G, env_int("LGPU_SLU_MU0", 4), env_int("LGPU_SLU_MU1", 2),
                             env_int("LGPU_SLU_TOP", 2));
    c->pairs.ensure(std::max<size_t>(c->splan.pair_records, 1) * PAIR_STRIDE);
    c->topfac.ensure(TOP_STRIDE);
    c->fwork.ensure(c->splan.work_rows * ROW_STRIDE);
    c->rhs.ensure(c->splan.rhs_vecs * SB);
    c->gvec.ensure(std::max<size_t>(c->splan.pair_records, 1) * SB);
    c->xpad.ensure(static_cast<size_t>(c->splan.n_pad) * BLK);
    {   // completion flags of the backward first stage: zero once, epochs only grow
      const size_t need = static_cast<size_t>(std::max(1, c->splan.stages[0].nchunks));
      if (need > c->solve_flags.cap) {
        c->solve_flags.ensure(need);
        CUDA_CHECK(cudaMemsetAsync(c->solve_flags.p, 0, need * sizeof(unsigned long long), c->stream));
      }
    }
    c->d_info.ensure(1);
    c->mbox.ensure(slu_mbox_elems(c->splan));
  }
  ensure_vectors(c);
  c->log.stream = c->stream;
  c->solve_epoch = 0;
  // every mailbox entry empty (all-ones NaN) before the first solve with these factors
  CUDA_CHECK(cudaMemsetAsync(c->mbox.p, 0xFF, sizeof(cd) * slu_mbox_elems(c->splan), c->stream));
  CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
  slu_factorize(c->splan, c->sdev(), sigma, c->stream, &c->log);
  CUDA_CHECK(cudaEventRecord(c->ev1, c->stream));
  int32_t info = 0;
  int32_t bmeta[2] = {c->bell_w, c->bell_real ? 0 : 1};   // longest row, any imaginary entry
  if (c->bell_w < 0) {   // B changed since the last factorisation: refresh its compressed copy
    const size_t rows = static_cast<size_t>(c->N);
    c->bell_val.ensure(rows * ELL_MAX_WIDTH);
    c->bell_col.ensure(rows * ELL_MAX_WIDTH);
    c->bell_width.ensure(2);
    c->bell_rval.ensure(rows * ELL_MAX_WIDTH);
    bell_build(c->G, c->B.p, BEll{c->bell_val.p, c->bell_col.p, c->bell_width.p, c->bell_rval.p}, c->stream, &c->log);
    CUDA_CHECK(cudaMemcpyAsync(bmeta, c->bell_width.p, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_CHECK(cudaMemcpyAsync(&info, c->d_info.p, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(stream_sync(c));
  c->bell_w = bmeta[0];
  c->bell_real = bmeta[1] == 0;
  float ms = 0.f;
  CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->t_factor = ms;
  c->sigma = sigma;
  c->lu_info = info;
  c->factorized = true;
  return LGPU_OK;
}

// The residual vector of the Arnoldi driver is allocated with room for the padding node of an odd grid, so that the
// solve kernels write the operator's output in place (slu_solve: x_padded).
bool out_has_pad(lgpu_ctx* c, const cd* p) {
  return p != nullptr && p == c->resid.p && c->resid.cap >= static_cast<size_t>(c->splan.n_pad) * BLK;
}

// x = M^-1 b on the device (b, x may alias), with optional iterative refinement
void dev_solve(lgpu_ctx* c, const cd* b, cd* x, int refine) {
  const cd* rhs = b;
  if (refine > 0 && b == x) {   // refinement needs the original right-hand side
    CUDA_CHECK(cudaMemcpyAsync(c->vr.p, b, sizeof(cd) * c->N, cudaMemcpyDeviceToDevice, c->stream));
    rhs = c->vr.p;
  }
  slu_solve(c->splan, c->sdev(), rhs, x, c->stream, &c->log, nullptr, out_has_pad(c, x));
  for (int it = 0; it < refine; ++it) {
    // e = M^-1 (b - (A - sigma B) x) ; x += e
    block_matvec(c->G, c->A.p, c->B.p, c->factor_of_B ? cd{0.0, 0.0} : cd{-1.0, 0.0},
                 c->factor_of_B ? cd{-1.0, 0.0} : c->sigma, x, rhs, c->ve.p, c->stream, &c->log);
    slu_solve(c->splan, c->sdev(), c->ve.p, c->ve.p, c->stream, &c->log);
    vec_axpby(c->N, cd{1.0, 0.0}, x, cd{1.0, 0.0}, c->ve.p, c->stream, &c->log);
  }
}

// general mode (smod_arpack_general.f08:88-91): y = B^-1 A x on the device (x, y may alias)
void dev_apply_op_general(lgpu_ctx* c, const cd* x, cd* y, int refine) {
  block_matvec(c->G, c->A.p, c->B.p, cd{1.0, 0.0}, cd{0.0, 0.0}, x, nullptr, c->vu.p, c->stream, &c->log);
  dev_solve(c, c->vu.p, y, refine);
}

// y = M^-1 B x on the device (x, y may alias)
void dev_apply_op(lgpu_ctx* c, const cd* x, cd* y, int refine) {
  if (c->factor_of_B) return dev_apply_op_general(c, x, y, refine);
  static const bool use_ell = [] { const char* e = std::getenv("LGPU_B_ELL"); return !(e && e[0] == '0'); }();
  static const bool fuse_bx = [] { const char* e = std::getenv("LGPU_BX_FUSE"); return !(e && e[0] == '0'); }();
  // B real with short rows (everything but Hall): the solve kernels form B x themselves, no product launch
  if (use_ell && fuse_bx && refine == 0 && c->bell_real && c->bell_w >= 0 && c->bell_w <= 8 && x != y &&
      c->splan.stages.size() >= 2) {
    RhsEll ell;
    ell.val = c->bell_rval.p; ell.col = c->bell_col.p; ell.x = x; ell.rows = c->G * BLK; ell.width = c->bell_w;
    c->log.fused_bx += 1;
    slu_solve(c->splan, c->sdev(), nullptr, y, c->stream, &c->log, &ell, out_has_pad(c, y));
    return;
  }
  if (use_ell && c->bell_w >= 0 && c->bell_w <= ELL_MAX_WIDTH)
    bell_matvec(c->G, BEll{c->bell_val.p, c->bell_col.p, c->bell_width.p, c->bell_rval.p}, c->bell_w, c->bell_real, x, c->vu.p, c->stream,
                &c->log);
  else
    block_matvec(c->G, c->A.p, c->B.p, cd{0.0, 0.0}, cd{1.0, 0.0}, x, nullptr, c->vu.p, c->stream,
                 &c->log);
  dev_solve(c, c->vu.p, y, refine);
}

class CudaKrylovOps final : public KrylovOps {
 public:
  CudaKrylovOps(lgpu_ctx* c, int ncv, int refine) : c_(c), ncv_(ncv), refine_(refine) {}

  void init_residual() override { dev_apply_op(c_, c_->resid.p, c_->resid.p, refine_); }

  void extend(int k, int m) override {
    const int n = c_->N;
    const KrylovWork kw = kwork(c_);
    cd* V = c_->V.p;
    cd* H = c_->Hdev.p;
    const BasisLayout& L = c_->basis;
    (void)n;
    if (k == 0) krylov_update(L, V, 0, c_->resid.p, kw, c_->stream, &c_->log);   // rnorm = ||resid||
    bool have_vj = false;   // did the previous step already normalise v_j into V(:, j) and vcur?
    for (int j = k; j < m; ++j) {
      if (!have_vj) {
        cd* hsub = j > 0 ? H + static_cast<size_t>(j - 1) * ncv_ + j : nullptr;
        krylov_scale(L, c_->resid.p, V, j, c_->vcur.p, kw, hsub, c_->stream, &c_->log);
      }
      std::unique_lock<std::mutex> dbg_step_lock;   // LGPU_DBG_SERIAL bit 3: one Arnoldi step of the process at a time
      if (dbg_serial_steps()) dbg_step_lock = std::unique_lock<std::mutex>(dbg_step_mutex());
      dev_apply_op(c_, c_->vcur.p, c_->resid.p, refine_);
      const bool dual = dbg_dual() && c_->dbg_cnt.p != nullptr;
      cd* dbg_win = nullptr; cd* dbg_wout = nullptr; cd* dbg_h = nullptr;
      if (dual) {
        dbg_win = c_->dbg_buf.p; dbg_wout = dbg_win + n; dbg_h = dbg_wout + n;
        dev_apply_op(c_, c_->vcur.p, dbg_win, refine_);                       // second evaluation of w = OP v
        dbg_compare(c_->resid.p, dbg_win, n, c_->dbg_cnt.p + 0, c_->stream);
      }
      KrylovWork kws = kw;   // w = resid comes with completion flags when the solve ended with its flagged kernel
      if (!dual && refine_ == 0 && c_->signal.last != 0 && c_->signal.last == c_->signal.epoch && c_->signal.mu >= 1 &&
          c_->signal.x == static_cast<const void*>(c_->resid.p)) {
        kws.wflags = c_->signal.flags; kws.wepoch = c_->signal.last;
        kws.wtile_shift = c_->signal.mu - 1; kws.wnchunks = c_->splan.stages[0].nchunks;
      }
      cd* hcol = H + static_cast<size_t>(j) * ncv_;
      // the fused step also produces v_{j+1} when there is a next step in this batch
      const int newcol = j + 1 < m ? j + 1 : -1;
      cd* hnext = H + static_cast<size_t>(j) * ncv_ + j + 1;
      have_vj = krylov_cgs2_step(L, V, j + 1, c_->resid.p, kws, hcol, newcol, c_->vcur.p, hnext, c_->stream,
                                 &c_->log);
      if (dual && have_vj) {   // the same step once more on the same w: outputs must agree bit for bit
        CUDA_CHECK(cudaMemcpyAsync(dbg_wout, c_->resid.p, sizeof(cd) * n, cudaMemcpyDeviceToDevice, c_->stream));
        CUDA_CHECK(cudaMemcpyAsync(dbg_h, hcol, sizeof(cd) * (j + 2), cudaMemcpyDeviceToDevice, c_->stream));
        CUDA_CHECK(cudaMemcpyAsync(c_->resid.p, dbg_win, sizeof(cd) * n, cudaMemcpyDeviceToDevice, c_->stream));
        krylov_cgs2_step(L, V, j + 1, c_->resid.p, kws, hcol, newcol, c_->vcur.p, hnext, c_->stream, &c_->log);
        dbg_compare(c_->resid.p, dbg_wout, n, c_->dbg_cnt.p + 1, c_->stream);
        dbg_compare(hcol, dbg_h, j + 2, c_->dbg_cnt.p + 2, c_->stream);
      }
      rnorm_in_h_ = have_vj && j + 1 == ncv_;   // the fused step left ||resid|| in the slot behind H (hnext of the last column)
      if (dbg_step_lock.owns_lock()) CUDA_CHECK(cudaStreamSynchronize(c_->stream));
      if (have_vj) {
        have_vj = newcol >= 0;
      } else {
        krylov_dots(L, V, j + 1, c_->resid.p, kw, hcol, 0, c_->stream, &c_->log);
        krylov_update_dots(L, V, j + 1, c_->resid.p, kw, hcol, 1, c_->stream, &c_->log);
        krylov_update(L, V, j + 1, c_->resid.p, kw, c_->stream, &c_->log);
      }
    }
  }

  void fetch(int k, int m, cplx* H, int ldh, double* rnorm) override {
    const size_t cnt = static_cast<size_t>(ncv_) * ncv_;
    c_->h_stage.ensure(cnt + 1);
    // one copy when the residual norm travels with H (the usual case), a second one for it otherwise
    const bool one = rnorm_in_h_ && m == ncv_;
    CUDA_CHECK(cudaMemcpyAsync(c_->h_stage.p, c_->Hdev.p, (cnt + (one ? 1 : 0)) * sizeof(cd), cudaMemcpyDeviceToHost,
                               c_->stream));
    if (!one)
      CUDA_CHECK(cudaMemcpyAsync(c_->h_scal.p, c_->kscal.p, sizeof(double), cudaMemcpyDeviceToHost, c_->stream));
    CUDA_CHECK(stream_sync(c_));
    if (one) c_->h_scal.p[0] = c_->h_stage.p[cnt].x;
    for (int j = k; j < m; ++j) {
      for (int i = 0; i <= j; ++i) {
        const cd v = c_->h_stage.p[static_cast<size_t>(j) * ncv_ + i];
        H[static_cast<size_t>(j) * ldh + i] = cplx(v.x, v.y);
      }
      if (j > 0) {
        const cd v = c_->h_stage.p[static_cast<size_t>(j - 1) * ncv_ + j];
        H[static_cast<size_t>(j - 1) * ldh + j] = cplx(v.x, v.y);
      }
    }
    *rnorm = c_->h_scal.p[0];
  }

  // Upload through a staging buffer of its own: the download staging buffer (fetch) is free to be
  // rewritten, and the copy enqueued here is complete by the time the host gets back (every
  // restart passes through the stream synchronisation of the next fetch()).
  void upload_small(const cplx* M, int ld, int rows, int cols) {
    const size_t cnt = static_cast<size_t>(ncv_) * ncv_;
    c_->h_upload.ensure(cnt);
    for (int j = 0; j < cols; ++j)
      for (int i = 0; i < rows; ++i) {
        const cplx v = M[static_cast<size_t>(j) * ld + i];
        c_->h_upload.p[static_cast<size_t>(j) * ncv_ + i] = cd{v.real(), v.imag()};
      }
    CUDA_CHECK(cudaMemcpyAsync(c_->Qdev.p, c_->h_upload.p, cnt * sizeof(cd), cudaMemcpyHostToDevice,
                               c_->stream));
  }

  void compress(int kplusp, int kev, const cplx* Q, int ldq, cplx sigmak, double betak) override {
    const int n = c_->N;
    upload_small(Q, ldq, kplusp, kev + 1 <= kplusp ? kev + 1 : kplusp);
    const int nc = kev + 1 <= kplusp ? kev + 1 : kplusp;
    (void)n;
    basis_gemm(c_->basis, c_->V.p, kplusp, c_->Qdev.p, ncv_, nc, c_->V.p, 0, c_->stream, &c_->log);
    vec_axpby_basis_norm(c_->basis, cd{sigmak.real(), sigmak.imag()}, c_->resid.p, cd{betak, 0.0},
                         c_->V.p, kev, kwork(c_), c_->stream, &c_->log);
  }

  void ritz_vectors(int kplusp, int nconv, const cplx* S, int lds) override {
    const int n = c_->N;
    upload_small(S, lds, kplusp, nconv);
    basis_gemm(c_->basis, c_->V.p, kplusp, c_->Qdev.p, ncv_, nconv, c_->Z.p, n, c_->stream,
               &c_->log);
  }

 private:
  lgpu_ctx* c_;
  int ncv_;
  int refine_;
  bool rnorm_in_h_ = false;
};

// general = false: OP = (A - sigma B)^-1 B, omega = sigma + 1/nu  (smod_arpack_shift_invert.f08)
// general = true:  OP = B^-1 A,            omega = nu              (smod_arpack_general.f08)
int do_shift_invert(lgpu_ctx* c, const lgpu_arnoldi* cfg, const double* resid0, bool resid_on_device,
                    double* omega_ri, double* vr_out, bool vr_on_device, lgpu_stats* stats,
                    bool general = false) {
  if (!c->assembled()) return fail(c, LGPU_ESTATE, "arnoldi: matrices not assembled");
  const int n = c->N;
  if (cfg->nev <= 0 || cfg->nev >= c->Nc()) return fail(c, LGPU_EINVAL, "nev out of range");
  if (cfg->ncv - cfg->nev < 1 || cfg->ncv > c->Nc() || cfg->ncv > KRYLOV_MAXCOL)
    return fail(c, LGPU_EINVAL, "ncv out of range (nev + 1 <= ncv <= min(N, 128))");
  if (cfg->maxiter <= 0) return fail(c, LGPU_EINVAL, "maxiter must be positive");
  static const char* allowed[6] = {"LM", "SM", "LR", "SR", "LI", "SI"};
  bool ok = false;
  for (auto w : allowed) ok = ok || (w[0] == cfg->which[0] && w[1] == cfg->which[1]);
  if (!ok) return fail(c, LGPU_EINVAL, "which must be one of LM SM LR SR LI SI");

  int rc = do_factorize(c, cd{cfg->sigma_re, cfg->sigma_im}, general);
  if (rc != LGPU_OK) return rc;
  DbgSerial dbg_serial(2);
  const int ncv = cfg->ncv, nev = cfg->nev;
  c->basis = make_basis_layout(n, ncv);
  c->V.ensure(c->basis.elems());
  // the fused Gram-Schmidt kernels stream whole groups of four columns and all 64 rows of the last
  // tile: entries outside the current basis must be finite (they meet zero coefficients)
  CUDA_CHECK(cudaMemsetAsync(c->V.p, 0, sizeof(cd) * c->basis.elems(), c->stream));
  c->vcur.ensure(n);
  c->resid.ensure(static_cast<size_t>(n) + BLK);   // + the padding node of an odd grid (out_has_pad)
  c->Hdev.ensure(static_cast<size_t>(ncv) * ncv + 1);   // + ||resid|| behind the last column (CudaKrylovOps::fetch)
  c->Qdev.ensure(static_cast<size_t>(ncv) * ncv);
  c->Z.ensure(static_cast<size_t>(n) * nev);
  ensure_krylov_work(c);
  CUDA_CHECK(cudaMemsetAsync(c->Hdev.p, 0, sizeof(cd) * ncv * ncv, c->stream));
  vec_in(c, resid0, resid_on_device, c->resid.p);
  if (dbg_dual()) {
    c->dbg_buf.ensure(2 * static_cast<size_t>(n) + BLK + ncv + 2);
    c->dbg_cnt.ensure(4);
    CUDA_CHECK(cudaMemsetAsync(c->dbg_cnt.p, 0, 4 * sizeof(unsigned long long), c->stream));
  }

  IramConfig ic;
  ic.nev = nev; ic.ncv = ncv; ic.maxiter = cfg->maxiter; ic.tol = cfg->tol;
  ic.which[0] = cfg->which[0]; ic.which[1] = cfg->which[1];
  CudaKrylovOps ops(c, ncv, cfg->refine_steps);
  Iram iram;
  const double t0 = now_ms();
  IramResult res = iram.run(ops, ic);
  CUDA_CHECK(stream_sync(c));
  if (dbg_dual()) {
    unsigned long long cnt[4] = {0, 0, 0, 0};
    CUDA_CHECK(cudaMemcpy(cnt, c->dbg_cnt.p, sizeof(cnt), cudaMemcpyDeviceToHost));
    if (cnt[0] | cnt[1] | cnt[2])
      std::fprintf(stderr, "[lgpu dual] ctx %p sigma (%g, %g): words differing between two evaluations: operator %llu, "
                   "step w %llu, step h %llu (n_op %d)\n", static_cast<void*>(c), cfg->sigma_re, cfg->sigma_im, cnt[0], cnt[1],
                   cnt[2], res.n_op);
  }
  const double t1 = now_ms();
  const double nan = std::numeric_limits<double>::quiet_NaN();
  for (int k = 0; k < nev; ++k) {
    if (k < res.nconv) {
      const cplx om = general ? res.ritz[k]
                              : cplx(cfg->sigma_re, cfg->sigma_im) + 1.0 / res.ritz[k];   // :157
      omega_ri[2 * k] = om.real();
      omega_ri[2 * k + 1] = om.imag();
    } else {
      omega_ri[2 * k] = nan;
      omega_ri[2 * k + 1] = nan;
    }
  }
  if (vr_out) {
    const size_t got = static_cast<size_t>(c->Nc()) * res.nconv, all = static_cast<size_t>(c->Nc()) * nev;
    if (vr_on_device) {
      if (res.nconv > 0) vec_out(c, c->Z.p, vr_out, true, res.nconv);
      if (all > got)
        CUDA_CHECK(cudaMemsetAsync(reinterpret_cast<cd*>(vr_out) + got, 0, (all - got) * sizeof(cd),
                                   c->stream));
      CUDA_CHECK(stream_sync(c));
    } else {
      if (res.nconv > 0) vec_out(c, c->Z.p, vr_out, false, res.nconv);
      CUDA_CHECK(stream_sync(c));
      if (all > got) std::memset(vr_out + 2 * got, 0, (all - got) * sizeof(cd));
    }
  }
  const double t2 = now_ms();
  c->t_iter = t1 - t0;
  c->t_extract = t2 - t1;
  if (general) {   // the factors of B are of no use to lgpu_solve / lgpu_apply_op
    c->factorized = false;
    c->factor_of_B = false;
  }
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->info = res.info;
    stats->nconv = res.nconv;
    stats->n_op = res.n_op;
    stats->n_bx = 0;
    stats->n_reorth = res.n_reorth;
    stats->n_restart = res.n_iter;
    stats->lu_info = c->lu_info;
    stats->t_factor_ms = c->t_factor;
    stats->t_iter_ms = c->t_iter;
    stats->t_extract_ms = c->t_extract;
  }
  return LGPU_OK;
}

}  // namespace

// ======================================================================= C entry points
extern "C" {

int lgpu_create(lgpu_ctx** out, int32_t device, int32_t log_level) {
  if (!out) return LGPU_EINVAL;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count)
    return LGPU_ENOGPU;
  std::unique_ptr<lgpu_ctx> c(new (std::nothrow) lgpu_ctx);
  if (!c) return LGPU_ENOMEM;
  c->device = device;
  c->log_level = log_level;
  if (cudaSetDevice(device) != cudaSuccess) return LGPU_ENOGPU;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return LGPU_ENOGPU;
  c->own_stream = true;
  if (cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess)
    return LGPU_ENOGPU;
  *out = c.release();
  return LGPU_OK;
}

int lgpu_destroy(lgpu_ctx* ctx) {
  if (!ctx) return LGPU_EINVAL;
  cudaSetDevice(ctx->device);
  stream_sync(ctx);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev_block) cudaEventDestroy(ctx->ev_block);
  if (ctx->sm_limit > 0) sharing_contexts() -= 1;
  cudaStream_t own = ctx->own_stream ? ctx->stream : nullptr;
  delete ctx;
  if (own) cudaStreamDestroy(own);
  return LGPU_OK;
}

const char* lgpu_last_error(const lgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int lgpu_set_stream(lgpu_ctx* ctx, void* cuda_stream) {
  return guarded(ctx, [&] {
    CUDA_CHECK(stream_sync(ctx));
    if (ctx->own_stream) {
      if (cuda_stream == nullptr) return LGPU_OK;
      cudaStreamDestroy(ctx->stream);
      ctx->own_stream = false;
    }
    if (cuda_stream == nullptr) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
      ctx->own_stream = true;
    } else {
      ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    }
    return LGPU_OK;
  });
}

int lgpu_set_sm_limit(lgpu_ctx* ctx, int32_t max_sms) {
  if (!ctx || max_sms < 0) return LGPU_EINVAL;
  if ((ctx->sm_limit > 0) != (max_sms > 0)) sharing_contexts() += max_sms > 0 ? 1 : -1;
  ctx->sm_limit = max_sms;
  ctx->demand_G = -1;
  return LGPU_OK;
}

int lgpu_synchronize(lgpu_ctx* ctx) {
  return guarded(ctx, [&] {
    CUDA_CHECK(stream_sync(ctx));
    return LGPU_OK;
  });
}

int lgpu_assemble(lgpu_ctx* ctx, const lgpu_settings* s, const double* base_grid,
                  const double* gauss_grid, const double* const fields[LGPU_N_FIELDS]) {
  return guarded(ctx, [&] {
    if (!s || !base_grid || !gauss_grid || !fields) return fail(ctx, LGPU_EINVAL, "null argument");
    if (s->gridpts < 2) return fail(ctx, LGPU_EINVAL, "gridpts must be >= 2");
    const size_t G = s->gridpts, ng = 4 * (G - 1);
    ctx->d_grid.ensure(G);
    ctx->d_gauss.ensure(ng);
    ctx->d_fields.ensure(ng * NFIELD);
    CUDA_CHECK(cudaMemcpyAsync(ctx->d_grid.p, base_grid, G * sizeof(double), cudaMemcpyHostToDevice,
                               ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(ctx->d_gauss.p, gauss_grid, ng * sizeof(double),
                               cudaMemcpyHostToDevice, ctx->stream));
    FieldPtrs fp{};
    for (int f = 0; f < NFIELD; ++f) {
      if (!fields[f]) continue;
      double* dst = ctx->d_fields.p + static_cast<size_t>(f) * ng;
      CUDA_CHECK(cudaMemcpyAsync(dst, fields[f], ng * sizeof(double), cudaMemcpyHostToDevice,
                                 ctx->stream));
      fp.f[f] = dst;
    }
    return do_assemble(ctx, s, ctx->d_grid.p, ctx->d_gauss.p, fp);
  });
}

int lgpu_assemble_device(lgpu_ctx* ctx, const lgpu_settings* s, const double* base_grid,
                         const double* gauss_grid, const double* const fields[LGPU_N_FIELDS]) {
  return guarded(ctx, [&] {
    if (!s || !base_grid || !gauss_grid || !fields) return fail(ctx, LGPU_EINVAL, "null argument");
    if (s->gridpts < 2) return fail(ctx, LGPU_EINVAL, "gridpts must be >= 2");
    FieldPtrs fp{};
    for (int f = 0; f < NFIELD; ++f) fp.f[f] = fields[f];
    return do_assemble(ctx, s, base_grid, gauss_grid, fp);
  });
}

int lgpu_matrix_dim(lgpu_ctx* ctx, int32_t* n) {
  if (!ctx || !n) return LGPU_EINVAL;
  *n = ctx->Nc();   // the reference's dim_matrix = gridpts * 2 * nb_eqs
  return LGPU_OK;
}

int lgpu_export_blocks(lgpu_ctx* ctx, int32_t which, double* blocks_ri) {
  return guarded(ctx, [&] {
    if (which < 0 || which > 1 || !blocks_ri) return fail(ctx, LGPU_EINVAL, "bad argument");
    if (!ctx->have[which]) return fail(ctx, LGPU_ESTATE, "matrix not available");
    const size_t cnt = static_cast<size_t>(ctx->G) * 3 * BLK2;
    if (ctx->compact()) {   // (G, 3, dsub, dsub) tiles in the reference's numbering
      std::vector<cd> blocks(cnt);
      CUDA_CHECK(cudaMemcpyAsync(blocks.data(), which ? ctx->B.p : ctx->A.p, cnt * sizeof(cd),
                                 cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_CHECK(stream_sync(ctx));
      const int d = ctx->dsub;
      for (size_t t = 0; t < static_cast<size_t>(ctx->G) * 3; ++t)
        for (int j = 0; j < d; ++j)
          for (int i = 0; i < d; ++i) {
            const cd v = blocks[t * BLK2 + ctx->cmap[j] * BLK + ctx->cmap[i]];
            double* dst = blocks_ri + 2 * (t * d * d + static_cast<size_t>(j) * d + i);
            dst[0] = v.x;
            dst[1] = v.y;
          }
      return LGPU_OK;
    }
    CUDA_CHECK(cudaMemcpyAsync(blocks_ri, which ? ctx->B.p : ctx->A.p, cnt * sizeof(cd),
                               cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(stream_sync(ctx));
    return LGPU_OK;
  });
}

int lgpu_export_coo(lgpu_ctx* ctx, int32_t which, int64_t* nnz, int32_t* rows, int32_t* cols,
                    double* vals_ri) {
  return guarded(ctx, [&] {
    if (which < 0 || which > 1 || !nnz) return fail(ctx, LGPU_EINVAL, "bad argument");
    if (!ctx->have[which]) return fail(ctx, LGPU_ESTATE, "matrix not available");
    const int G = ctx->G;
    const size_t cnt = static_cast<size_t>(G) * 3 * BLK2;
    std::vector<cd> blocks(cnt);
    std::vector<uint32_t> masks(static_cast<size_t>(G) * MASK_WORDS), nat(2 * 2 * 4 * 8 + 2);
    CUDA_CHECK(cudaMemcpyAsync(blocks.data(), which ? ctx->B.p : ctx->A.p, cnt * sizeof(cd),
                               cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(masks.data(), ctx->masks.p, masks.size() * sizeof(uint32_t),
                               cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(nat.data(), ctx->natmasks.p, nat.size() * sizeof(uint32_t),
                               cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(stream_sync(ctx));
    auto bit = [](const uint32_t* w, int idx) { return (w[idx >> 5] >> (idx & 31)) & 1u; };
    int inv[BLK];   // device row within a block -> row in the reference's numbering
    for (int i = 0; i < BLK; ++i) inv[i] = 0;
    for (int q = 0; q < ctx->dsub; ++q) inv[ctx->cmap[q]] = q;
    int64_t count = 0;
    const bool fill = rows && cols && vals_ri;
    for (int b = 0; b < G; ++b) {
      const uint32_t* mw = masks.data() + static_cast<size_t>(b) * MASK_WORDS + which * 32;
      for (int i = 0; i < BLK; ++i) {
        // per-row insertion order of the reference's linked list: element b-1 (sub, diag),
        // element b (new diag entries, super), natural boundary, essential diagonal
        int ent_tile[3 * BLK + 1], ent_col[3 * BLK + 1], ne = 0;
        bool present[3][BLK] = {};
        auto add = [&](int tile, int j) {
          if (present[tile][j]) return;
          present[tile][j] = true;
          ent_tile[ne] = tile; ent_col[ne] = j; ++ne;
        };
        for (int j = 0; j < BLK; ++j) if (bit(mw + 0 * 8, j * BLK + i)) add(0, j);
        for (int j = 0; j < BLK; ++j) if (bit(mw + 1 * 8, j * BLK + i)) add(1, j);
        for (int j = 0; j < BLK; ++j) if (bit(mw + 2 * 8, j * BLK + i)) add(1, j);
        for (int j = 0; j < BLK; ++j) if (bit(mw + 3 * 8, j * BLK + i)) add(2, j);
        for (int edge = 0; edge < 2; ++edge) {
          const int qr = edge == 0 ? b : b - (G - 2);
          if (qr < 0 || qr > 1) continue;
          for (int qc = 0; qc < 2; ++qc) {
            const uint32_t* nw = nat.data() + ((edge * 2 + which) * 4 + qr * 2 + qc) * 8;
            for (int j = 0; j < BLK; ++j)
              if (bit(nw, j * BLK + i)) add(qc - qr + 1, j);
          }
          if (which == 1 && ((nat[2 * 2 * 4 * 8 + edge] >> (qr * BLK + i)) & 1u)) add(1, i);
        }
        for (int e = 0; e < ne; ++e) {
          if (fill) {
            const cd v = blocks[(static_cast<size_t>(b) * 3 + ent_tile[e]) * BLK2 + ent_col[e] * BLK + i];
            rows[count] = b * ctx->dsub + inv[i] + 1;
            cols[count] = (b + ent_tile[e] - 1) * ctx->dsub + inv[ent_col[e]] + 1;
            vals_ri[2 * count] = v.x;
            vals_ri[2 * count + 1] = v.y;
          }
          ++count;
        }
      }
    }
    *nnz = count;
    return LGPU_OK;
  });
}

int lgpu_import_coo(lgpu_ctx* ctx, int32_t which, int32_t n, int64_t nnz, const int32_t* rows,
                    const int32_t* cols, const double* vals_ri) {
  return guarded(ctx, [&] {
    if (which < 0 || which > 1 || n <= 0 || n % BLK != 0 || nnz < 0)
      return fail(ctx, LGPU_EINVAL, "import_coo: n must be a positive multiple of 16");
    const int G = n / BLK;
    if (ctx->have[1 - which] && ctx->G != G) ctx->have[1 - which] = false;   // new problem size
    const size_t cnt = static_cast<size_t>(G) * 3 * BLK2;
    std::vector<cd> blocks(cnt, cd{0.0, 0.0});
    std::vector<uint32_t> masks(static_cast<size_t>(G) * MASK_WORDS, 0u);
    if (ctx->have[1 - which] && ctx->masks.p) {
      CUDA_CHECK(cudaMemcpy(masks.data(), ctx->masks.p, masks.size() * sizeof(uint32_t),
                            cudaMemcpyDeviceToHost));
      for (int b = 0; b < G; ++b)
        for (int wd = 0; wd < 32; ++wd) masks[static_cast<size_t>(b) * MASK_WORDS + which * 32 + wd] = 0u;
    }
    for (int64_t k = 0; k < nnz; ++k) {
      const int r = rows[k] - 1, c = cols[k] - 1;
      if (r < 0 || r >= n || c < 0 || c >= n) return fail(ctx, LGPU_EINVAL, "import_coo: index out of range");
      const int b = r / BLK, t = c / BLK - b + 1;
      if (t < 0 || t > 2) return fail(ctx, LGPU_EINVAL, "import_coo: entry outside the block-tridiagonal envelope");
      const int idx = (c % BLK) * BLK + r % BLK;
      cd& dst = blocks[(static_cast<size_t>(b) * 3 + t) * BLK2 + idx];
      dst.x += vals_ri[2 * k];
      dst.y += vals_ri[2 * k + 1];
      const int contrib = t == 0 ? 0 : (t == 1 ? 1 : 3);
      masks[static_cast<size_t>(b) * MASK_WORDS + which * 32 + contrib * 8 + (idx >> 5)] |= 1u << (idx & 31);
    }
    ctx->G = G;
    ctx->N = n;
    ctx->dsub = BLK;   // imported matrices are in the full 16-wide numbering
    for (int i = 0; i < BLK; ++i) ctx->cmap[i] = i;
    ctx->padmask = 0;
    ctx->factorized = false;
    if (ctx->splan.n != G) ctx->splan.n = 0;
    (which ? ctx->B : ctx->A).ensure(cnt);
    ctx->masks.ensure(masks.size());
    ctx->natmasks.ensure(2 * 2 * 4 * 8 + 2);
    CUDA_CHECK(cudaMemcpy((which ? ctx->B : ctx->A).p, blocks.data(), cnt * sizeof(cd), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(ctx->masks.p, masks.data(), masks.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemset(ctx->natmasks.p, 0, (2 * 2 * 4 * 8 + 2) * sizeof(uint32_t)));
    ctx->have[which] = true;
    ctx->have_grid = false;
    if (which == 1) ctx->bell_w = -1;
    return LGPU_OK;
  });
}

int lgpu_factorize(lgpu_ctx* ctx, double sigma_re, double sigma_im, int32_t* lu_info) {
  return guarded(ctx, [&] {
    const int rc = do_factorize(ctx, cd{sigma_re, sigma_im});
    if (rc == LGPU_OK && lu_info) *lu_info = ctx->lu_info;
    return rc;
  });
}

int lgpu_solve(lgpu_ctx* ctx, const double* rhs_ri, double* x_ri, int32_t refine_steps) {
  return guarded(ctx, [&] {
    if (!ctx->factorized) return fail(ctx, LGPU_ESTATE, "solve: call lgpu_factorize first");
    if (!rhs_ri || !x_ri) return fail(ctx, LGPU_EINVAL, "null argument");
    vec_in(ctx, rhs_ri, false, ctx->vx.p);
    dev_solve(ctx, ctx->vx.p, ctx->vy.p, refine_steps);
    vec_out(ctx, ctx->vy.p, x_ri, false);
    CUDA_CHECK(stream_sync(ctx));
    return LGPU_OK;
  });
}

int lgpu_matvec(lgpu_ctx* ctx, int32_t which, const double* x_ri, double* y_ri) {
  return guarded(ctx, [&] {
    if (which < 0 || which > 1 || !x_ri || !y_ri) return fail(ctx, LGPU_EINVAL, "bad argument");
    if (!ctx->assembled()) return fail(ctx, LGPU_ESTATE, "matvec: matrices not assembled");
    ensure_vectors(ctx);
    vec_in(ctx, x_ri, false, ctx->vx.p);
    block_matvec(ctx->G, ctx->A.p, ctx->B.p, cd{which == 0 ? 1.0 : 0.0, 0.0},
                 cd{which == 1 ? 1.0 : 0.0, 0.0}, ctx->vx.p, nullptr, ctx->vy.p, ctx->stream,
                 &ctx->log);
    vec_out(ctx, ctx->vy.p, y_ri, false);
    CUDA_CHECK(stream_sync(ctx));
    return LGPU_OK;
  });
}

int lgpu_apply_op(lgpu_ctx* ctx, const double* x_ri, double* y_ri, int32_t refine_steps) {
  return guarded(ctx, [&] {
    if (!ctx->factorized) return fail(ctx, LGPU_ESTATE, "apply_op: call lgpu_factorize first");
    if (!x_ri || !y_ri) return fail(ctx, LGPU_EINVAL, "null argument");
    vec_in(ctx, x_ri, false, ctx->vx.p);
    dev_apply_op(ctx, ctx->vx.p, ctx->vy.p, refine_steps);
    vec_out(ctx, ctx->vy.p, y_ri, false);
    CUDA_CHECK(stream_sync(ctx));
    return LGPU_OK;
  });
}

int lgpu_apply_op_device(lgpu_ctx* ctx, const double* x_dev, double* y_dev, int32_t refine_steps, int32_t repeat,
                         double* ms_per_application) {
  return guarded(ctx, [&] {
    if (!ctx->factorized) return fail(ctx, LGPU_ESTATE, "apply_op_device: call lgpu_factorize first");
    if (!x_dev || !y_dev || x_dev == y_dev || repeat < 1) return fail(ctx, LGPU_EINVAL, "apply_op_device: bad argument");
    if (ctx->compact()) return fail(ctx, LGPU_EINVAL, "apply_op_device: mhd state vector only (device vectors are in the 16-wide layout)");
    const cd* x = reinterpret_cast<const cd*>(x_dev);
    cd* y = reinterpret_cast<cd*>(y_dev);
    // as the Arnoldi driver issues it: from its current-vector buffer into its (padded) residual buffer
    const size_t n = static_cast<size_t>(ctx->N);
    ctx->vcur.ensure(n);
    ctx->resid.ensure(n + BLK);
    CUDA_CHECK(cudaMemcpyAsync(ctx->vcur.p, x, n * sizeof(cd), cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int r = 0; r < repeat; ++r) dev_apply_op(ctx, ctx->vcur.p, ctx->resid.p, refine_steps);
    CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(y, ctx->resid.p, n * sizeof(cd), cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_CHECK(cudaEventSynchronize(ctx->ev1));
    CUDA_CHECK(stream_sync(ctx));
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (ms_per_application) *ms_per_application = static_cast<double>(ms) / repeat;
    return LGPU_OK;
  });
}

namespace {

// hwork[0..2] of the last vec_dot2 -> host (synchronises the stream)
void fetch_dots(lgpu_ctx* c, cd out[3]) {
  c->h_stage.ensure(4);
  CUDA_CHECK(cudaMemcpyAsync(c->h_stage.p, c->khwork.p, 3 * sizeof(cd), cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(stream_sync(c));
  for (int k = 0; k < 3; ++k) out[k] = c->h_stage.p[k];
}

void dev_bx(lgpu_ctx* c, const cd* x, cd* y) {
  if (c->bell_w >= 0 && c->bell_w <= ELL_MAX_WIDTH)
    bell_matvec(c->G, BEll{c->bell_val.p, c->bell_col.p, c->bell_width.p, c->bell_rval.p}, c->bell_w, c->bell_real, x, y, c->stream, &c->log);
  else
    block_matvec(c->G, c->A.p, c->B.p, cd{0.0, 0.0}, cd{1.0, 0.0}, x, nullptr, y, c->stream, &c->log);
}

}  // namespace

int lgpu_residuals(lgpu_ctx* ctx, int32_t nev, const double* omega_ri, const double* vr_ri, double* res) {
  return guarded(ctx, [&] {
    if (!ctx->assembled()) return fail(ctx, LGPU_ESTATE, "residuals: matrices not assembled");
    if (nev < 0 || (nev > 0 && (!omega_ri || !vr_ri || !res))) return fail(ctx, LGPU_EINVAL, "null argument");
    ensure_vectors(ctx);
    ensure_krylov_work(ctx);
    ctx->log.stream = ctx->stream;
    const KrylovWork kw = kwork(ctx);
    for (int k = 0; k < nev; ++k) {
      const cd om{omega_ri[2 * k], omega_ri[2 * k + 1]};
      if (std::fabs(om.x) <= DP_LIMIT && std::fabs(om.y) <= DP_LIMIT) { res[k] = 0.0; continue; }   // is_zero
      vec_in(ctx, vr_ri + 2 * static_cast<size_t>(k) * ctx->Nc(), false, ctx->vx.p);
      // y = A v - omega B v
      block_matvec(ctx->G, ctx->A.p, ctx->B.p, cd{1.0, 0.0}, cd{-om.x, -om.y}, ctx->vx.p, nullptr, ctx->vy.p,
                   ctx->stream, &ctx->log);
      cd d[3], e[3];
      vec_dot2(ctx->N, ctx->vy.p, ctx->vy.p, ctx->vy.p, kw, ctx->stream, &ctx->log);
      fetch_dots(ctx, d);
      vec_dot2(ctx->N, ctx->vx.p, ctx->vx.p, ctx->vx.p, kw, ctx->stream, &ctx->log);
      fetch_dots(ctx, e);
      res[k] = std::sqrt(d[2].x) / (std::hypot(om.x, om.y) * std::sqrt(e[2].x));
    }
    return LGPU_OK;
  });
}

int lgpu_eigenfunctions(lgpu_ctx* ctx, const double* vr_ri, int32_t nsel, const int32_t* idxs, double* out_ri) {
  return guarded(ctx, [&] {
    if (!ctx->assembled() || !ctx->have_grid)
      return fail(ctx, LGPU_ESTATE, "eigenfunctions: needs matrices assembled by lgpu_assemble (grid, geometry)");
    if (nsel < 0 || (nsel > 0 && (!vr_ri || !idxs || !out_ri))) return fail(ctx, LGPU_EINVAL, "null argument");
    if (nsel == 0) return LGPU_OK;
    const size_t n = static_cast<size_t>(ctx->N);
    const int npts = 2 * ctx->G - 1;
    ctx->ef_in.ensure(n * nsel);
    ctx->ef_out.ensure(static_cast<size_t>(8) * npts * nsel);
    ctx->ef_idx.ensure(nsel);
    std::vector<int32_t> local(nsel);
    const size_t nc = static_cast<size_t>(ctx->Nc());
    for (int s = 0; s < nsel; ++s) {
      if (idxs[s] < 1) return fail(ctx, LGPU_EINVAL, "eigenfunctions: indices are 1-based");
      local[s] = s;   // the selected columns are packed on the way to the device
      vec_in(ctx, vr_ri + 2 * nc * static_cast<size_t>(idxs[s] - 1), false, ctx->ef_in.p + n * s);
    }
    CUDA_CHECK(cudaMemcpyAsync(ctx->ef_idx.p, local.data(), sizeof(int32_t) * nsel, cudaMemcpyHostToDevice, ctx->stream));
    ctx->log.stream = ctx->stream;
    assemble_eigenfunctions(ctx->G, ctx->settings.geometry, ctx->grid_copy.p, ctx->ef_in.p, n, nsel, ctx->ef_idx.p,
                            ctx->ef_out.p, ctx->stream, &ctx->log);
    // one slab per variable of the active state vector, in state-vector order
    const size_t slab = static_cast<size_t>(npts) * nsel;
    for (int q = 0; q < ctx->dsub / 2; ++q)
      CUDA_CHECK(cudaMemcpyAsync(out_ri + 2 * slab * q, ctx->ef_out.p + slab * (ctx->cmap[2 * q] / 2), sizeof(cd) * slab,
                                 cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(stream_sync(ctx));
    return LGPU_OK;
  });
}

int lgpu_inverse_iteration(lgpu_ctx* ctx, double sigma_re, double sigma_im, int32_t maxiter, double tol,
                           double* omega_ri, double* vr_ri, lgpu_stats* stats) {
  return guarded(ctx, [&] {
    if (!ctx->assembled()) return fail(ctx, LGPU_ESTATE, "inverse_iteration: matrices not assembled");
    if (!omega_ri) return fail(ctx, LGPU_EINVAL, "null argument");
    if (maxiter == 0) maxiter = 100;                 // smod_inverse_iteration.f08:59-61
    if (maxiter < 0) return fail(ctx, LGPU_EINVAL, "maxiter has to be positive");
    if (sigma_re == 0.0 && sigma_im == 0.0) return fail(ctx, LGPU_EINVAL, "inverse-iteration: sigma can not be equal to zero");
    int rc = do_factorize(ctx, cd{sigma_re, sigma_im});
    if (rc != LGPU_OK) return rc;
    ensure_krylov_work(ctx);
    const KrylovWork kw = kwork(ctx);
    lgpu_ctx* c = ctx;
    const int n = c->N;
    cd* x = c->vx.p;
    cd* r = c->vr.p;
    cd* sv = c->vy.p;
    cd d[3];
    auto normalise = [&](cd* v) {
      vec_dot2(n, v, v, v, kw, c->stream, &c->log);
      fetch_dots(c, d);
      const double inv = 1.0 / std::sqrt(d[2].x);
      vec_axpby(n, cd{inv, 0.0}, v, cd{0.0, 0.0}, v, c->stream, &c->log);
    };
    // start vector: (A - sigma B)^-1 1
    {
      std::vector<cd> ones(static_cast<size_t>(c->Nc()), cd{1.0, 0.0});
      vec_in(c, ones.data(), false, x);
      CUDA_CHECK(stream_sync(c));
    }
    slu_solve(c->splan, c->sdev(), x, x, c->stream, &c->log);
    normalise(x);
    int i = 0;
    bool converged = false;
    cd ev{sigma_re, sigma_im};
    while (i <= maxiter && !converged) {
      dev_bx(c, x, r);                                                                        // r = B x
      block_matvec(c->G, c->A.p, c->B.p, cd{1.0, 0.0}, cd{0.0, 0.0}, x, nullptr, sv, c->stream, &c->log);   // s = A x
      vec_dot2(n, x, sv, r, kw, c->stream, &c->log);
      fetch_dots(c, d);
      ev = d[0] * crecip(d[1]);                                                                // x^H s / x^H r
      vec_axpby(n, cd{1.0, 0.0}, sv, cd{-ev.x, -ev.y}, r, c->stream, &c->log);                 // s -= ev r
      vec_dot2(n, sv, sv, sv, kw, c->stream, &c->log);
      cd e[3];
      fetch_dots(c, e);
      if (std::sqrt(e[2].x) < std::hypot(ev.x, ev.y) * tol) { converged = true; break; }
      ++i;
      slu_solve(c->splan, c->sdev(), r, x, c->stream, &c->log);                                // x = M^-1 r
      normalise(x);
    }
    omega_ri[0] = ev.x;
    omega_ri[1] = ev.y;
    if (vr_ri) {
      std::vector<cd> h(static_cast<size_t>(c->Nc()));
      vec_out(c, x, h.data(), false);
      CUDA_CHECK(stream_sync(c));
      // make the largest coefficient real (first maximum, as idamax)
      size_t im = 0;
      double best = -1.0;
      for (size_t k = 0; k < h.size(); ++k) {
        const double a = std::hypot(h[k].x, h[k].y);
        if (a > best) { best = a; im = k; }
      }
      const cd ph{h[im].x / best, -h[im].y / best};
      for (size_t k = 0; k < h.size(); ++k) {
        const cd v = h[k] * ph;
        vr_ri[2 * k] = v.x;
        vr_ri[2 * k + 1] = v.y;
      }
    }
    if (stats) {
      *stats = lgpu_stats{};
      stats->info = converged ? 0 : 1;
      stats->nconv = converged ? 1 : 0;
      stats->n_op = i;
      stats->lu_info = c->lu_info;
      stats->t_factor_ms = c->t_factor;
    }
    return LGPU_OK;
  });
}

int lgpu_shift_invert(lgpu_ctx* ctx, const lgpu_arnoldi* cfg, const double* resid0_ri,
                      double* omega_ri, double* vr_ri, lgpu_stats* stats) {
  return guarded(ctx, [&] {
    if (!cfg || !resid0_ri || !omega_ri) return fail(ctx, LGPU_EINVAL, "null argument");
    return do_shift_invert(ctx, cfg, resid0_ri, false, omega_ri, vr_ri, false, stats);
  });
}

int lgpu_arnoldi_general(lgpu_ctx* ctx, const lgpu_arnoldi* cfg, const double* resid0_ri,
                         double* omega_ri, double* vr_ri, lgpu_stats* stats) {
  return guarded(ctx, [&] {
    if (!cfg || !resid0_ri || !omega_ri) return fail(ctx, LGPU_EINVAL, "arnoldi_general: null argument");
    return do_shift_invert(ctx, cfg, resid0_ri, false, omega_ri, vr_ri, false, stats, true);
  });
}

int lgpu_shift_invert_device(lgpu_ctx* ctx, const lgpu_arnoldi* cfg, const double* resid0_dev,
                             double* omega_ri_host, double* vr_dev, lgpu_stats* stats) {
  return guarded(ctx, [&] {
    if (!cfg || !resid0_dev || !omega_ri_host) return fail(ctx, LGPU_EINVAL, "null argument");
    return do_shift_invert(ctx, cfg, resid0_dev, true, omega_ri_host, vr_dev, true, stats);
  });
}

void* lgpu_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes > 0 ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void lgpu_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}

int lgpu_zlarnv(int32_t iseed[4], int32_t n, double* out_ri) {
  if (!iseed || n < 0 || (n > 0 && !out_ri)) return LGPU_EINVAL;
  // LAPACK dlaruv: x_i = seed * a^i mod 2^48, a = 33952834046453, 128 numbers per call;
  // zlarnv(idist = 2) draws 64 complex numbers per dlaruv call and maps u -> 2u - 1.
  const uint64_t a = 33952834046453ull, mask = (1ull << 48) - 1;
  uint64_t s = (static_cast<uint64_t>(iseed[0]) << 36) | (static_cast<uint64_t>(iseed[1]) << 24) |
               (static_cast<uint64_t>(iseed[2]) << 12) | static_cast<uint64_t>(iseed[3]);
  const int64_t total = 2 * static_cast<int64_t>(n);
  const double scale = 1.0 / 281474976710656.0;   // 2^-48
  for (int64_t pos = 0; pos < total;) {
    const int64_t cnt = std::min<int64_t>(128, total - pos);
    uint64_t m = 1;
    for (int64_t i = 0; i < cnt; ++i) {
      m = (m * a) & mask;
      const uint64_t x = (s * m) & mask;
      out_ri[pos + i] = 2.0 * (static_cast<double>(x) * scale) - 1.0;
    }
    s = (s * m) & mask;
    pos += cnt;
  }
  iseed[0] = static_cast<int32_t>((s >> 36) & 4095);
  iseed[1] = static_cast<int32_t>((s >> 24) & 4095);
  iseed[2] = static_cast<int32_t>((s >> 12) & 4095);
  iseed[3] = static_cast<int32_t>(s & 4095);
  return LGPU_OK;
}

int lgpu_counters(lgpu_ctx* ctx, int64_t* kernel_launches, int32_t reset) {
  if (!ctx) return LGPU_EINVAL;
  if (kernel_launches) *kernel_launches = ctx->log.launches;
  if (reset) ctx->log.launches = 0;
  return LGPU_OK;
}

int lgpu_set_profiling(lgpu_ctx* ctx, int32_t enable) {
  return guarded(ctx, [&] {
    CUDA_CHECK(stream_sync(ctx));
    ctx->log.stream = ctx->stream;
    ctx->log.reset();
    ctx->log.profiling = enable != 0;
    return LGPU_OK;
  });
}

int lgpu_profile_read(lgpu_ctx* ctx, double* ms, int64_t* counts, double* algo_bytes,
                      int32_t nkinds, int32_t reset) {
  return guarded(ctx, [&] {
    CUDA_CHECK(stream_sync(ctx));
    ctx->log.collect();
    for (int k = 0; k < nkinds && k < LK_COUNT; ++k) {
      if (ms) ms[k] = ctx->log.ms[k];
      if (counts) counts[k] = ctx->log.count[k];
      if (algo_bytes) algo_bytes[k] = ctx->log.bytes[k];
    }
    if (reset) ctx->log.reset();
    return LGPU_OK;
  });
}

int lgpu_phase_times(lgpu_ctx* ctx, double* t_assemble_ms, double* t_factor_ms, double* t_iter_ms,
                     double* t_extract_ms) {
  if (!ctx) return LGPU_EINVAL;
  if (t_assemble_ms) *t_assemble_ms = ctx->t_assemble;
  if (t_factor_ms) *t_factor_ms = ctx->t_factor;
  if (t_iter_ms) *t_iter_ms = ctx->t_iter;
  if (t_extract_ms) *t_extract_ms = ctx->t_extract;
  return LGPU_OK;
}

}  // extern "C"
