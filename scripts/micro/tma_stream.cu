// Microbenchmark: per-SM streaming rate of cp.async.bulk (global -> shared) through a ring of stages,
// one CTA per SM, versus the number of stages and the copy size; and a plain LDG.128 streaming read.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  do { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.b32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory"); } while (!ok);
}
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}
// nprod producer lanes split each stage's copy
__global__ void __launch_bounds__(288, 1) tma_kernel(const char* src, size_t per_cta, int bytes, int ns, int nprod, int interleave, double* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)ns * bytes);
  uint64_t* empty = full + ns;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { for (int i = 0; i < ns; ++i) { mb_init(&full[i], 1); mb_init(&empty[i], 8); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const int ntile = (int)(per_cta / bytes);
  if (warp == 8) {
    if (lane < nprod) {
      int st = 0; uint32_t ph = 0;
      for (int u = 0; u < ntile; ++u) {
        if (u >= ns) mb_wait(&empty[st], ph ^ 1u);
        const size_t off = interleave ? ((size_t)u * gridDim.x + blockIdx.x) * bytes : (size_t)blockIdx.x * per_cta + (size_t)u * bytes;
        if (lane == 0) mb_expect(&full[st], bytes);
        __syncwarp((1u << nprod) - 1);
        const int part = bytes / nprod;
        bulk(smem + (size_t)st * bytes + lane * part, src + off + lane * part, part, &full[st]);
        if (++st == ns) { st = 0; ph ^= 1u; }
      }
    }
  } else {
    int st = 0; uint32_t ph = 0; double acc = 0;
    for (int u = 0; u < ntile; ++u) {
      mb_wait(&full[st], ph);
      acc += reinterpret_cast<const double*>(smem + (size_t)st * bytes)[tid];
      __syncwarp();
      if (lane == 0) mb_arrive(&empty[st]);
      if (++st == ns) { st = 0; ph ^= 1u; }
    }
    if (acc == 1.2345e300) sink[tid] = acc;
  }
}
__global__ void __launch_bounds__(1024, 1) ldg_kernel(const double2* src, size_t n16, double* sink) {
  double acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) { const double2 v = __ldg(src + i); acc += v.x + v.y; }
  if (acc == 1.2345e300) sink[threadIdx.x] = acc;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t total = (size_t)sms * 16 * 1024 * 1024;   // 16 MiB per SM = 2.3 GiB: far beyond L2
  char* src; cudaMalloc(&src, total); cudaMemset(src, 0, total);
  double* sink; cudaMalloc(&sink, 8192);
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](int bytes, int ns, int nprod, int il, size_t per_cta) {
    const size_t smem = (size_t)ns * bytes + 16 * ns;
    tma_kernel<<<sms, 288, smem>>>(src, per_cta, bytes, ns, nprod, il, sink); cudaDeviceSynchronize();
    cudaEventRecord(e0); tma_kernel<<<sms, 288, smem>>>(src, per_cta, bytes, ns, nprod, il, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double b = (double)(per_cta / bytes) * bytes * sms;
    printf("bulk %6d B x %2d stages, %2d lanes, %s, %5.1f MB total: %7.1f us  %6.0f GB/s (%5.1f GB/s per SM) err=%d\n", bytes, ns, nprod, il ? "interleaved" : "per-CTA blocks", b / 1e6, ms * 1e3, b / ms * 1e-6, b / ms * 1e-6 / sms, (int)cudaGetLastError());
  };
  const size_t big = 16 * 1024 * 1024, small = 512 * 1024;   // per CTA: 2.3 GiB total / 76 MB total
  for (size_t pc : {big, small}) {
    run(32768, 4, 1, 0, pc); run(32768, 4, 1, 1, pc); run(32768, 6, 1, 1, pc); run(16384, 8, 1, 1, pc); run(8192, 16, 1, 1, pc);
    run(32768, 4, 4, 1, pc); run(32768, 4, 16, 1, pc); run(65536, 3, 1, 1, pc); run(4096, 32, 1, 1, pc);
  }
  for (size_t n : {total, (size_t)sms * small}) {
    ldg_kernel<<<sms * 2, 1024>>>((const double2*)src, n / 16, sink); cudaDeviceSynchronize();
    cudaEventRecord(e0); ldg_kernel<<<sms * 2, 1024>>>((const double2*)src, n / 16, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("LDG.128 streaming read %7.1f MB: %7.1f us %6.0f GB/s\n", n / 1e6, ms * 1e3, n / ms * 1e-6);
  }
  return 0;
}
