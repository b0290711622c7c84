// Shared helpers: complex128 value type with mixed real/complex operators, CUDA error
// macros, and the device-side constants of the hot path.
#pragma once

#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace lgpu {

// complex(dp): 16-byte aligned so a whole value moves with one 128-bit load/store.
struct __align__(16) cd {
  double x, y;
};

__host__ __device__ inline cd mk(double x, double y = 0.0) { return cd{x, y}; }
__host__ __device__ inline cd operator+(cd a, cd b) { return cd{a.x + b.x, a.y + b.y}; }
__host__ __device__ inline cd operator-(cd a, cd b) { return cd{a.x - b.x, a.y - b.y}; }
__host__ __device__ inline cd operator-(cd a) { return cd{-a.x, -a.y}; }
__host__ __device__ inline cd operator*(cd a, cd b) {
  return cd{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
__host__ __device__ inline cd operator+(cd a, double b) { return cd{a.x + b, a.y}; }
__host__ __device__ inline cd operator+(double a, cd b) { return cd{a + b.x, b.y}; }
__host__ __device__ inline cd operator-(cd a, double b) { return cd{a.x - b, a.y}; }
__host__ __device__ inline cd operator-(double a, cd b) { return cd{a - b.x, -b.y}; }
__host__ __device__ inline cd operator*(cd a, double b) { return cd{a.x * b, a.y * b}; }
__host__ __device__ inline cd operator*(double a, cd b) { return cd{a * b.x, a * b.y}; }
__host__ __device__ inline cd operator/(cd a, double b) { return cd{a.x / b, a.y / b}; }
__host__ __device__ inline cd& operator+=(cd& a, cd b) { a.x += b.x; a.y += b.y; return a; }
__host__ __device__ inline cd& operator-=(cd& a, cd b) { a.x -= b.x; a.y -= b.y; return a; }
__host__ __device__ inline cd conj(cd a) { return cd{a.x, -a.y}; }
__host__ __device__ inline double abs2(cd a) { return a.x * a.x + a.y * a.y; }
// a += b*c (4 FMAs)
__host__ __device__ inline void cfma(cd& a, cd b, cd c) {
  a.x = fma(b.x, c.x, a.x); a.x = fma(-b.y, c.y, a.x);
  a.y = fma(b.x, c.y, a.y); a.y = fma(b.y, c.x, a.y);
}
// a -= b*c
__host__ __device__ inline void cfms(cd& a, cd b, cd c) {
  a.x = fma(-b.x, c.x, a.x); a.x = fma(b.y, c.y, a.x);
  a.y = fma(-b.x, c.y, a.y); a.y = fma(-b.y, c.x, a.y);
}
// a += conj(b)*c
__host__ __device__ inline void cfmac(cd& a, cd b, cd c) {
  a.x = fma(b.x, c.x, a.x); a.x = fma(b.y, c.y, a.x);
  a.y = fma(b.x, c.y, a.y); a.y = fma(-b.y, c.x, a.y);
}
// complex reciprocal (Smith's algorithm, like Fortran's complex division)
__host__ __device__ inline cd crecip(cd a) {
  if (fabs(a.x) >= fabs(a.y)) {
    double r = a.y / a.x, den = a.x + a.y * r;
    return cd{1.0 / den, -r / den};
  }
  double r = a.x / a.y, den = a.x * r + a.y;
  return cd{r / den, -1.0 / den};
}

constexpr int BLK = 16;             // dim_subblock for MHD (src/settings/mod_dims.f08:36-44)
constexpr int BLK2 = BLK * BLK;     // entries per block
constexpr double DP_LIMIT = 5.0e-15;  // src/mod_global_variables.f08:19

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t err, const char* what, const char* file, int line) {
  if (err != cudaSuccess) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(err));
    throw CudaError(buf);
  }
}
#define CUDA_CHECK(expr) ::lgpu::cuda_check((expr), #expr, __FILE__, __LINE__)

// Function attributes (dynamic shared-memory opt-in) and the SM count are per DEVICE: a process may
// hold contexts on several GPUs (lgpu_create(device = n)), so one-time configuration is keyed by the
// current device, not by a process-wide flag.
struct PerDeviceOnce {
  std::mutex m;
  uint64_t done = 0;
  template <typename F>
  void run(F&& f) {
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> g(m);
    if (dev < 64 && ((done >> dev) & 1ull)) return;
    f();
    if (dev < 64) done |= 1ull << dev;
  }
};
inline int device_sm_count() {
  static std::mutex m;
  static int cache[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> g(m);
  if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  if (dev >= 0 && dev < 64) cache[dev] = n;
  return n;
}

// Kernel-launch accounting and optional per-kernel-class device timing (CUDA events recorded
// on the launching stream around each launch; read back at a synchronisation point).
enum LaunchKind {
  LK_ASSEMBLE = 0, LK_FACTOR, LK_MATVEC, LK_FWD0, LK_FWD, LK_TOP, LK_BWD, LK_BWD0, LK_DOTS,
  LK_UPDATE, LK_SCALE, LK_GEMM, LK_CGS2, LK_OTHER, LK_COUNT
};

// ---- mbarrier + bulk asynchronous copy (TMA without a tensor map), sm_90+ PTX ----------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// Release of a ring slot that a bulk copy (async proxy) will refill, by a warp that has read it with ordinary
// (generic-proxy) loads.  The mbarrier orders the arrival against the producer lane's wait, not the two proxies against
// each other: a warp can arrive while its loads of the slot are still outstanding, and without the cross-proxy fence
// the refill could land first - the warp then worked on the NEXT record.  Never seen with one context on the GPU, but
// with CTAs of other contexts on the same SMs one operator application in ~10^3 came out wrong
// (profiles/tuning_log_r2.md, "several contexts in flight").  Every lane fences its own loads, one lane arrives.
// (-DLGPU_FENCE_AT_PRODUCER: the fence in front of every copy instead, issued by the producer lane after it has
// acquired the slot - equally effective, but it keeps the copies of a ring from overlapping: 45 -> 58 us per
// Gram-Schmidt step.)
__device__ __forceinline__ void mbar_release_slot(uint64_t* empty, int lane) {
#ifndef LGPU_FENCE_AT_PRODUCER
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
  __syncwarp();
  if (lane == 0) mbar_arrive(empty);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t addr = smem_u32(b);
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                 " selp.b32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
#ifdef LGPU_FENCE_AT_PRODUCER
  asm volatile("fence.proxy.async;" ::: "memory");
#endif
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// L2 eviction-priority policies for bulk copies (the stream access-policy window does not apply
// to the async proxy): evict_first for data streamed once per launch, evict_last for the small
// hot set that every launch re-reads.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
#ifdef LGPU_DBG_SAFE_TMA   // debugging variant: no L2 hint
  bulk_g2s(dst, src, bytes, bar);
  return;
#endif
#ifdef LGPU_FENCE_AT_PRODUCER
  asm volatile("fence.proxy.async;" ::: "memory");
#endif
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
#endif   // __CUDACC__

struct LaunchLog {
  int64_t launches = 0;
  int64_t fused_bx = 0;            // operator applications whose B x product ran inside the solve kernels
  bool profiling = false;
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> pool;   // start/stop pairs
  std::vector<int> kinds;          // kind of each recorded pair
  size_t used = 0;                 // events used
  double ms[LK_COUNT] = {};
  int64_t count[LK_COUNT] = {};
  double bytes[LK_COUNT] = {};     // algorithmic bytes of the timed launches

  void begin(int kind, double algo_bytes = 0.0) {
    if (!profiling) return;
    bytes[kind] += algo_bytes;
    if (used + 2 > pool.size()) {
      const size_t grow = pool.size() ? pool.size() : 4096;
      for (size_t i = 0; i < grow; ++i) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) { profiling = false; return; }
        pool.push_back(e);
      }
    }
    kinds.push_back(kind);
    cudaEventRecord(pool[used], stream);
  }
  void end() {
    if (!profiling || kinds.size() * 2 <= used) return;
    cudaEventRecord(pool[used + 1], stream);
    used += 2;
  }
  // requires the stream to be idle
  void collect() {
    for (size_t i = 0; i + 1 < used; i += 2) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, pool[i], pool[i + 1]) == cudaSuccess) {
        ms[kinds[i / 2]] += t;
        count[kinds[i / 2]] += 1;
      }
    }
    used = 0;
    kinds.clear();
  }
  void reset() {
    collect();
    for (int k = 0; k < LK_COUNT; ++k) { ms[k] = 0.0; count[k] = 0; bytes[k] = 0.0; }
  }
  ~LaunchLog() { for (auto e : pool) cudaEventDestroy(e); }
};

}  // namespace lgpu
