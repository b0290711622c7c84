// Microbenchmark: kernel-to-kernel gap on one stream with and without programmatic dependent launch
// (PDL), for plain and cooperative launches.  Answers: (1) can cudaLaunchAttributeCooperative and
// cudaLaunchAttributeProgrammaticStreamSerialization be combined, (2) what a dependent launch costs
// when the secondary's prologue may overlap the primary's tail.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_coop pdl_coop.cu && ./pdl_coop
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
namespace cg = cooperative_groups;

__global__ void work_kernel(long long spin, int pdl, unsigned long long* sink) {
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  const long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0 && blockIdx.x == 0 && spin < 0) *sink = t0;
}

__global__ void coop_kernel(long long spin, int pdl, unsigned long long* sink) {
  cg::grid_group g = cg::this_grid();
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  const long long t0 = clock64();
  while (clock64() - t0 < spin / 2) {}
  g.sync();
  while (clock64() - t0 < spin) {}
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0 && blockIdx.x == 0 && spin < 0) *sink = t0;
}

static cudaError_t launch(void* fn, int grid, bool coop, bool pdl, cudaStream_t s, long long spin, unsigned long long* sink) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (coop) { at[na].id = cudaLaunchAttributeCooperative; at[na].val.cooperative = 1; ++na; }
  if (pdl) { at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
  cfg.attrs = at; cfg.numAttrs = na;
  int p = pdl ? 1 : 0;
  void* args[] = {&spin, &p, &sink};
  return cudaLaunchKernelExC(&cfg, fn, args);
}

int main() {
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  unsigned long long* sink; cudaMalloc(&sink, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const long long spin = 20000;   // ~10 us
  const int reps = 200;
  for (int coop = 0; coop < 2; ++coop)
    for (int pdl = 0; pdl < 2; ++pdl) {
      cudaError_t err = cudaSuccess;
      for (int w = 0; w < 2 && err == cudaSuccess; ++w) {
        cudaEventRecord(e0, s);
        for (int i = 0; i < reps && err == cudaSuccess; ++i) {
          err = launch((void*)work_kernel, 296, false, pdl, s, spin, sink);
          if (err == cudaSuccess) err = launch(coop ? (void*)coop_kernel : (void*)work_kernel, 148, coop, pdl, s, spin, sink);
        }
        cudaEventRecord(e1, s);
        cudaError_t e2 = cudaStreamSynchronize(s);
        if (err == cudaSuccess) err = e2;
      }
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      printf("second kernel %s, pdl=%d: %s, %.2f us per kernel (busy %.2f us)\n", coop ? "cooperative" : "plain", pdl,
             cudaGetErrorString(err), 1e3 * ms / (2 * reps), spin / 1.965e3);
      cudaGetLastError();
    }
  return 0;
}
