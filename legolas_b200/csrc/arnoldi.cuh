// K3b / K3c — device kernels of the Arnoldi process: classical Gram-Schmidt with
// unconditional re-orthogonalisation (CGS2) as tall-skinny complex GEMVs, basis
// normalisation, and the tall-skinny GEMM used by the implicit restart (V <- V Q) and the
// Ritz-vector extraction (Z = V S).
//
// Replaces what ARPACK's znaitr / znapps / zneupd do with zgemv / zgemm on the host
// (arpack-ng, called from src/solvers/arnoldi/smod_arpack_shift_invert.f08:64-81,119-143).
#pragma once

#include <cstdint>

#include "common.cuh"

namespace lgpu {

constexpr int KRYLOV_TILE = 512;     // rows per CTA in the dot / update kernels
constexpr int KRYLOV_MAXCOL = 128;   // ncv limit of the device kernels

struct KrylovWork {
  cd* partial;        // [max_ctas][KRYLOV_MAXCOL + 1]
  cd* hwork;          // [KRYLOV_MAXCOL + 1] coefficients of the current projection
  double* scal;       // [0] rnorm  [1] scratch
  unsigned int* ticket;
};

// h = V(:, 0:ncols)^H w  -> work.hwork ; Hcol (device, may be null) gets `=` (accumulate == 0)
// or `+=` (accumulate == 1).
void krylov_dots(int n, const cd* V, int ldv, int ncols, const cd* w, const KrylovWork& work,
                 cd* Hcol, int accumulate, cudaStream_t stream, LaunchLog* log);
// w -= V(:, 0:ncols) hwork ; afterwards scal[0] = ||w||_2
void krylov_update(int n, const cd* V, int ldv, int ncols, cd* w, const KrylovWork& work,
                   cudaStream_t stream, LaunchLog* log);
// scal[0] = ||w||_2
void krylov_norm(int n, const cd* w, const KrylovWork& work, cudaStream_t stream,
                 LaunchLog* log);
// vout = w / scal[0] ; if hsub != null: *hsub = (scal[0], 0)
void krylov_scale(int n, const cd* w, cd* vout, const KrylovWork& work, cd* hsub,
                  cudaStream_t stream, LaunchLog* log);
// Out(:, 0:nc) = V(:, 0:nk) Q(0:nk, 0:nc); Q is a device matrix with leading dimension ldq.
// Out may alias V (row-local update).
void basis_gemm(int n, const cd* V, int ldv, int nk, const cd* Q, int ldq, int nc, cd* Out,
                int ldo, cudaStream_t stream, LaunchLog* log);
// r = a*r + b*v
void vec_axpby(int n, cd a, cd* r, cd b, const cd* v, cudaStream_t stream, LaunchLog* log);

}  // namespace lgpu
