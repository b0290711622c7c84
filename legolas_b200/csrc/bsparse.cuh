// K2'' — y = B x from a compressed copy of B.
//
// The reference applies B through zgbmv on the full band (src/solvers/arnoldi/mod_linear_systems /
// smod_arpack_shift_invert.f08: 63 diagonals of 16-byte entries per row).  The assembled B is far
// sparser than its band: the mass matrix couples each variable only with itself (<= 6 entries
// per row, more only with Hall / electron inertia).  Once per assembly the dense block rows of B
// are compacted on the device into a fixed-width ELL layout (entries in ascending column order,
// width = the longest row); every operator application then streams 20 bytes per stored entry
// instead of 768 bytes per row.  Rows longer than ELL_MAX_WIDTH keep the dense block kernel.
#pragma once

#include "common.cuh"

namespace lgpu {

constexpr int ELL_MAX_WIDTH = 16;

struct BEll {
  cd* val;         // [ELL_MAX_WIDTH][rows]
  int32_t* col;    // [ELL_MAX_WIDTH][rows]
  int32_t* width;  // device, 2 ints: [0] longest row found by the last build, [1] != 0 if any entry of
                   // B has an imaginary part (only with Hall / electron inertia, smod_hall_matrix.f08:6-86)
  double* rval;    // [ELL_MAX_WIDTH][rows] real parts: the product streams 12 instead of 20 bytes per
                   // entry when B is real (the reference itself writes B as a real matrix, mod_output.f08:494)
};

// Compact the (n, 3, 16, 16) block rows of B; *ell.width = max entries per row (may exceed
// ELL_MAX_WIDTH, in which case the ELL copy is unusable).
void bell_build(int n, const cd* B, const BEll& ell, cudaStream_t stream, LaunchLog* log);
// y = B x using `width` entries per row; real_only: every stored entry has a zero imaginary part
void bell_matvec(int n, const BEll& ell, int width, bool real_only, const cd* x, cd* y, cudaStream_t stream,
                 LaunchLog* log);

}  // namespace lgpu
