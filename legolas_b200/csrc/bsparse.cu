// K2'' — see bsparse.cuh.
#include "bsparse.cuh"

namespace lgpu {
namespace {

// one thread per matrix row: scan its 48 block entries in ascending column order
__global__ void __launch_bounds__(256) bell_build_kernel(int n, const cd* __restrict__ B, BEll ell) {
  const int rows = n * BLK;
  const int row = blockIdx.x * 256 + threadIdx.x;
  int count = 0, imag = 0;
  if (row < rows) {
    const int b = row / BLK, i = row - b * BLK;
    for (int t = 0; t < 3; ++t) {
      const int bc = b - 1 + t;
      if (bc < 0 || bc >= n) continue;
      const cd* blk = B + (static_cast<size_t>(b) * 3 + t) * BLK2;
      for (int j = 0; j < BLK; ++j) {
        const cd v = blk[j * BLK + i];
        if (v.x != 0.0 || v.y != 0.0) {
          if (count < ELL_MAX_WIDTH) {
            ell.val[static_cast<size_t>(count) * rows + row] = v;
            ell.rval[static_cast<size_t>(count) * rows + row] = v.x;
            ell.col[static_cast<size_t>(count) * rows + row] = bc * BLK + j;
          }
          if (v.y != 0.0) imag = 1;
          ++count;
        }
      }
    }
    for (int k = count; k < ELL_MAX_WIDTH; ++k) {
      ell.val[static_cast<size_t>(k) * rows + row] = cd{0.0, 0.0};
      ell.rval[static_cast<size_t>(k) * rows + row] = 0.0;
      ell.col[static_cast<size_t>(k) * rows + row] = row;
    }
  }
  // longest row of the CTA, then of the grid
  for (int off = 16; off >= 1; off >>= 1) count = max(count, __shfl_xor_sync(0xffffffffu, count, off));
  if ((threadIdx.x & 31) == 0) atomicMax(ell.width, count);
  if (__any_sync(0xffffffffu, imag) && (threadIdx.x & 31) == 0) atomicOr(ell.width + 1, 1);
}

template <int W>
__global__ void __launch_bounds__(256) bell_matvec_kernel(int rows, const cd* __restrict__ val,
                                                          const int32_t* __restrict__ col,
                                                          const cd* __restrict__ x, cd* __restrict__ y) {
  const int row = blockIdx.x * 256 + threadIdx.x;
  if (row >= rows) return;
  cd v[W];
  int32_t c[W];
#pragma unroll
  for (int k = 0; k < W; ++k) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(val + static_cast<size_t>(k) * rows + row));
    v[k] = cd{t.x, t.y};
    c[k] = __ldg(col + static_cast<size_t>(k) * rows + row);
  }
  cd acc{0.0, 0.0};
#pragma unroll
  for (int k = 0; k < W; ++k) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(x + c[k]));
    cfma(acc, v[k], cd{t.x, t.y});
  }
  y[row] = acc;
}

// B real: 8-byte values
template <int W>
__global__ void __launch_bounds__(256) bell_matvec_real_kernel(int rows, const double* __restrict__ val,
                                                               const int32_t* __restrict__ col,
                                                               const cd* __restrict__ x, cd* __restrict__ y) {
  const int row = blockIdx.x * 256 + threadIdx.x;
  if (row >= rows) return;
  double v[W];
  int32_t c[W];
#pragma unroll
  for (int k = 0; k < W; ++k) {
    v[k] = __ldg(val + static_cast<size_t>(k) * rows + row);
    c[k] = __ldg(col + static_cast<size_t>(k) * rows + row);
  }
  cd acc{0.0, 0.0};
#pragma unroll
  for (int k = 0; k < W; ++k) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(x + c[k]));
    acc.x = fma(v[k], t.x, acc.x);
    acc.y = fma(v[k], t.y, acc.y);
  }
  y[row] = acc;
}

}  // namespace

void bell_build(int n, const cd* B, const BEll& ell, cudaStream_t stream, LaunchLog* log) {
  CUDA_CHECK(cudaMemsetAsync(ell.width, 0, 2 * sizeof(int32_t), stream));
  const int rows = n * BLK;
  log->begin(LK_OTHER, 768.0 * rows + 20.0 * ELL_MAX_WIDTH * rows);
  bell_build_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(n, B, ell);
  log->end();
  log->launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

void bell_matvec(int n, const BEll& ell, int width, bool real_only, const cd* x, cd* y, cudaStream_t stream,
                 LaunchLog* log) {
  const int rows = n * BLK;
  const int grid = (rows + 255) / 256;
  // algorithmic bytes as for the band product it replaces (12288 B of B per block row + x, y)
  log->begin(LK_MATVEC, (12288.0 + 512.0) * n);
  if (real_only && width <= 8) {
    if (width <= 4) bell_matvec_real_kernel<4><<<grid, 256, 0, stream>>>(rows, ell.rval, ell.col, x, y);
    else if (width <= 6) bell_matvec_real_kernel<6><<<grid, 256, 0, stream>>>(rows, ell.rval, ell.col, x, y);
    else bell_matvec_real_kernel<8><<<grid, 256, 0, stream>>>(rows, ell.rval, ell.col, x, y);
  } else if (width <= 4) bell_matvec_kernel<4><<<grid, 256, 0, stream>>>(rows, ell.val, ell.col, x, y);
  else if (width <= 6) bell_matvec_kernel<6><<<grid, 256, 0, stream>>>(rows, ell.val, ell.col, x, y);
  else if (width <= 8) bell_matvec_kernel<8><<<grid, 256, 0, stream>>>(rows, ell.val, ell.col, x, y);
  else if (width <= 12) bell_matvec_kernel<12><<<grid, 256, 0, stream>>>(rows, ell.val, ell.col, x, y);
  else bell_matvec_kernel<16><<<grid, 256, 0, stream>>>(rows, ell.val, ell.col, x, y);
  log->end();
  log->launches += 1;
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace lgpu
