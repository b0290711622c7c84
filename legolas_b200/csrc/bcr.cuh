// K2 / K2' / K3a — block cyclic reduction (BCR) factorisation and solve of the 16-wide
// block-tridiagonal matrix  M = A - sigma*B,  and the block-tridiagonal matvec.
//
// Replaces (reference call sites):
//   A - sigma*B + zgbtrf .. src/solvers/arnoldi/smod_arpack_shift_invert.f08:56-59,
//                            src/solvers/mod_linear_systems.f08:102-127
//   zgbtrs ................. src/solvers/mod_linear_systems.f08:67-97
//   zgbmv .................. src/matrices/datastructure/mod_banded_operations.f08:18-41
//
// Elimination order: classic odd-even cyclic reduction.  Level l works on the rows whose
// index is a multiple of 2^l; the odd multiples are eliminated, the even ones kept.
// Per eliminated row j (neighbours j-s, j+s, s = 2^l) the factorisation stores
//     Dinv_j = D_j^-1,  GL_j = D_j^-1 L_j,  GU_j = D_j^-1 U_j        (pivoted Gauss-Jordan)
// and per kept row i the pre-update couplings (Lk_i, Uk_i).  A solve is then
//     forward   y_j = Dinv_j r_j ;  r_i -= Lk_i y_{i-s} + Uk_i y_{i+s}
//     backward  x_j = y_j - GL_j x_{j-s} - GU_j x_{j+s}
// i.e. 5 blocks (20 KB) streamed per eliminated row and solve — the same bytes LAPACK's
// pivoted band LU needs (94 x 16 complex per column), but with log2(G) dependent levels
// instead of 16*G dependent columns.
//
// Solve scheduling ("stages"): consecutive levels are fused into one kernel by giving each
// CTA a chunk of C = 2^m consecutive rows of the stage's base level; the chunk interior
// (C - 1 rows) is eliminated locally in m levels with the right-hand side in shared memory,
// only the two separator rows talk to neighbouring chunks (through delta vectors, so the
// result is deterministic: no floating-point atomics anywhere).
#pragma once

#include <cstdint>
#include <vector>

#include "common.cuh"

namespace lgpu {

struct BcrLevel {
  int n_active;      // rows taking part at this level
  int n_elim;        // floor(n_active / 2)
  int n_kept;        // ceil(n_active / 2)
  size_t off_dinv;   // factor storage offsets, in units of one 16x16 block
  size_t off_lkuk;   // (Lk, Uk) pairs, 2 blocks per kept row
  size_t off_glgu;   // (GL, GU) pairs, 2 blocks per eliminated row
  size_t off_work;   // (L, D, U) of the rows active at this level (factorisation only)
};

struct BcrStage {
  int l0;            // first global level of the stage
  int m;             // levels fused in the stage (chunk = 2^m rows of level l0)
  int n0;            // rows active at level l0
  int nchunks;
  size_t off_rin;    // compact right-hand side of the stage (in units of 16 complex)
  size_t off_delta;  // dL / dR of the stage, 2 * (nchunks + 1) entries
};

struct BcrPlan {
  int n = 0;                       // block rows
  std::vector<BcrLevel> levels;    // levels[l], l = 0 .. nlevels-1 (n_active >= 2)
  size_t off_root = 0;             // Dinv of the last remaining row
  size_t factor_blocks = 0;        // total factor storage
  size_t work_blocks = 0;          // factorisation workspace (levels >= 1)
  std::vector<BcrStage> stages;    // forward order; the last one is the single-CTA top stage
  size_t rhs_vecs = 0;             // compact rhs storage (16-complex units)
  size_t delta_vecs = 0;
};

// chunk_log2[s] = levels fused in stage s (the remaining levels go to the top stage).
BcrPlan make_bcr_plan(int n, int first_stage_m, int next_stage_m, int top_max_rows);

struct BcrDevice {
  const cd* A;          // (n, 3, 256) blocks
  const cd* B;
  cd* factors;          // plan.factor_blocks * 256
  cd* work;             // plan.work_blocks * 256
  cd* rhs;              // plan.rhs_vecs * 16
  cd* delta;            // plan.delta_vecs * 16
  cd* yvec;             // n * 16
  int32_t* info;        // singular-pivot report (1-based block row, 0 = none)
};

// factorise A - sigma*B; ~2 launches per level
void bcr_factorize(const BcrPlan& plan, const BcrDevice& d, cd sigma, cudaStream_t stream,
                   int64_t* launches);
// x = M^-1 b ; b and x are device vectors of n*16 complex (may alias)
void bcr_solve(const BcrPlan& plan, const BcrDevice& d, const cd* b, cd* x, cudaStream_t stream,
               int64_t* launches);
// y = (alpha_a*A + alpha_b*B) x (+ beta*z): covers B*x, A*x and the refinement residual
// r = b - (A - sigma*B) x  (alpha_a = -1, alpha_b = sigma, z = b)
void block_matvec(int n, const cd* A, const cd* B, cd alpha_a, cd alpha_b, const cd* x,
                  const cd* z, cd* y, cudaStream_t stream, int64_t* launches);

}  // namespace lgpu
