// K3b / K3c — device kernels of the Arnoldi process: classical Gram-Schmidt with
// unconditional re-orthogonalisation (CGS2) as tall-skinny complex GEMVs, basis
// normalisation, and the tall-skinny GEMM used by the implicit restart (V <- V Q) and the
// Ritz-vector extraction (Z = V S).
//
// Replaces what ARPACK's znaitr / znapps / zneupd do with zgemv / zgemm on the host
// (arpack-ng, called from src/solvers/arnoldi/smod_arpack_shift_invert.f08:64-81,119-143).
//
// Basis layout.  ARPACK keeps V column-major (N x ncv).  On HBM that makes every CTA of a
// tall-skinny kernel read ncv short streams 16*N bytes apart (hundreds of concurrent DRAM
// streams, ~25 % of the copy bandwidth measured).  Here V is tiled by rows: tile t holds
// rows [t*T, (t+1)*T) of ALL columns contiguously,
//       V(i, c)  at  ((i / T) * ncv + c) * T + i % T ,
// with T = 64: the first j columns of a tile are ONE contiguous block of j KB (a single bulk
// copy), and a CTA that owns a range of consecutive tiles streams one contiguous region of HBM.
// Only the library sees this layout: start vector, operator input and Ritz vectors are plain
// contiguous vectors.
#pragma once

#include <cstdint>

#include "common.cuh"

namespace lgpu {

constexpr int KRYLOV_MAXCOL = 128;   // ncv limit of the device kernels
constexpr int KRYLOV_TILE = 64;          // rows per tile
constexpr int KRYLOV_PASS_MAXCOL = 64;   // columns one Gram-Schmidt pass launch covers

struct BasisLayout {
  int n;        // rows
  int ncv;      // columns stored per tile
  int T;        // rows per tile (multiple of 32)
  int ntiles;
  size_t elems() const { return static_cast<size_t>(ntiles) * ncv * T; }
};

BasisLayout make_basis_layout(int n, int ncv);

struct KrylovWork {
  cd* partial;        // [ntiles][KRYLOV_MAXCOL + 1]
  cd* hwork;          // [KRYLOV_MAXCOL + 1] coefficients of the current projection
  double* scal;       // [0] rnorm  [1] scratch
  unsigned int* ticket;
  unsigned long long* gbar;         // device-wide barrier counter of the fused step kernel (or null)
  unsigned long long* gbar_count;   // host: value the counter will have when the next launch starts
  int grid_cap = 0;                 // > 0: the fused step kernel uses at most this many CTAs (lgpu_set_sm_limit)
  // w was produced in place by a solve whose last kernel publishes per-chunk completion flags (slu.cuh: SolveSignal):
  // rows [512 c .. ) << ... of w are final once wflags[c] >= wepoch.  Null: wait for the kernel before the step.
  const unsigned long long* wflags = nullptr;
  unsigned long long wepoch = 0;
  int wtile_shift = 0;              // chunk of tile t = min(t >> wtile_shift, wnchunks - 1)
  int wnchunks = 0;
};
// CTAs of the fused step kernel (one per SM, all of them must be resident: it has a device-wide barrier)
int krylov_cgs2_grid(int ntiles, int grid_cap);

// h = V(:, 0:ncols)^H w  -> work.hwork ; Hcol (device, may be null) gets `=` (accumulate == 0)
// or `+=` (accumulate == 1).
void krylov_dots(const BasisLayout& L, const cd* V, int ncols, const cd* w, const KrylovWork& work,
                 cd* Hcol, int accumulate, cudaStream_t stream, LaunchLog* log);
// w -= V(:, 0:ncols) hwork, then hwork = V(:, 0:ncols)^H w in the same pass over V (the middle
// pass of CGS2); Hcol as in krylov_dots
void krylov_update_dots(const BasisLayout& L, const cd* V, int ncols, cd* w, const KrylovWork& work,
                        cd* Hcol, int accumulate, cudaStream_t stream, LaunchLog* log);
// w -= V(:, 0:ncols) hwork ; afterwards scal[0] = ||w||_2   (ncols == 0: just the norm)
void krylov_update(const BasisLayout& L, const cd* V, int ncols, cd* w, const KrylovWork& work,
                   cudaStream_t stream, LaunchLog* log);
// One whole CGS2 step in a single cooperative launch: Hcol = h + s with h = V^H w, w -= V h,
// s = V^H w, w -= V s ; scal[0] = ||w|| ; if newcol >= 0 also V(:, newcol) = vplain = w / ||w|| and
// *hsub = ||w||.  Returns false (nothing launched) when the step does not fit this path (basis
// wider than KRYLOV_PASS_MAXCOL columns, more rows than one resident wave can keep on chip):
// the caller then runs the three pass kernels and krylov_scale.
bool krylov_cgs2_step(const BasisLayout& L, cd* V, int ncols, cd* w, const KrylovWork& work, cd* Hcol,
                      int newcol, cd* vplain, cd* hsub, cudaStream_t stream, LaunchLog* log);
// V(:, col) = vplain = w / scal[0] ; if hsub != null: *hsub = (scal[0], 0)
void krylov_scale(const BasisLayout& L, const cd* w, cd* V, int col, cd* vplain,
                  const KrylovWork& work, cd* hsub, cudaStream_t stream, LaunchLog* log);
// Out(:, 0:nc) = V(:, 0:nk) Q(0:nk, 0:nc); Q is a device matrix with leading dimension ldq.
// out_plain_ld == 0: Out is a tiled basis (may alias V: the update is row-local);
// otherwise Out is plain column-major with that leading dimension.
void basis_gemm(const BasisLayout& L, const cd* V, int nk, const cd* Q, int ldq, int nc, cd* Out,
                int out_plain_ld, cudaStream_t stream, LaunchLog* log);
// r = a*r + b*V(:, col)
void vec_axpby_basis(const BasisLayout& L, cd a, cd* r, cd b, const cd* V, int col,
                     cudaStream_t stream, LaunchLog* log);
// r = a*r + b*V(:, col), then scal[0] = ||r||_2 (deterministic), one launch
void vec_axpby_basis_norm(const BasisLayout& L, cd a, cd* r, cd b, const cd* V, int col, const KrylovWork& work,
                          cudaStream_t stream, LaunchLog* log);
// work.hwork[0] = x^H s, [1] = x^H r, [2] = (||x||^2, 0)   (plain vectors; deterministic)
void vec_dot2(int n, const cd* x, const cd* sv, const cd* rv, const KrylovWork& work, cudaStream_t stream,
              LaunchLog* log);
// r = a*r + b*v (plain vectors)
void vec_axpby(int n, cd a, cd* r, cd b, const cd* v, cudaStream_t stream, LaunchLog* log);

}  // namespace lgpu
