// Microbenchmark: dependent-chain latency of DFMA / DADD / SHFL / LDS.128 and DFMA throughput at low occupancy.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void dfma_chain(double* out, long long* cyc, int iters) {
  double a[ILP];
  for (int k = 0; k < ILP; ++k) a[k] = threadIdx.x * 1e-3 + k;
  const double b = 1.0000001, c = 1e-9;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) a[k] = fma(a[k], b, c);
  }
  const long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < ILP; ++k) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void shfl_chain(double* out, long long* cyc, int iters) {
  double a = threadIdx.x;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) a = __shfl_xor_sync(0xffffffffu, a, 1) + 1.0;
  const long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void lds_chain(double* out, long long* cyc, int iters) {
  __shared__ double2 buf[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_double2((i * 7 + 1) % 1024, 0.0);
  __syncthreads();
  int idx = threadIdx.x;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) idx = (int)buf[idx & 1023].x;
  const long long t1 = clock64();
  out[threadIdx.x] = idx;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMallocManaged(&cyc, 8);
  const int iters = 10000;
  for (int warps : {1, 8, 16, 32}) {
    dfma_chain<1><<<1, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    printf("warps/SM %2d ILP1: %.1f cyc/iter | ", warps, double(*cyc) / iters);
    dfma_chain<2><<<1, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    printf("ILP2: %.1f | ", double(*cyc) / iters);
    dfma_chain<4><<<1, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    printf("ILP4: %.1f | ", double(*cyc) / iters);
    dfma_chain<8><<<1, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    printf("ILP8: %.1f cyc/iter\n", double(*cyc) / iters);
  }
  shfl_chain<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
  printf("shfl(double)+dadd chain: %.1f cyc/iter\n", double(*cyc) / iters);
  lds_chain<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
  printf("LDS.128 + cvt dependent chain: %.1f cyc/iter\n", double(*cyc) / iters);
  return 0;
}
