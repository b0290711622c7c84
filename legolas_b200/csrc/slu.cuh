// K2 / K2' / K3a — pivoted block cyclic reduction ("structured LU") factorisation and solve of
// the 16-wide block-tridiagonal matrix  M = A - sigma*B,  and the block-tridiagonal matvec.
//
// Replaces (reference call sites):
//   A - sigma*B + zgbtrf .. src/solvers/arnoldi/smod_arpack_shift_invert.f08:56-59,
//                            src/solvers/mod_linear_systems.f08:102-127
//   zgbtrs ................. src/solvers/mod_linear_systems.f08:67-97
//   zgbmv .................. src/matrices/datastructure/mod_banded_operations.f08:18-41
//
// Why not plain block cyclic reduction: the MHD operator is dominated by first-order
// derivative terms whose diagonal-block contributions cancel (centred-difference structure),
// so the 16x16 diagonal blocks of A - sigma*B are O(sigma*dx) while the off-diagonal blocks
// are O(1).  Pivoting restricted to the diagonal block then has growth ~1/(sigma*dx) per
// level and the solve loses all accuracy at 10^4 grid points (measured).
// zgbtrf avoids this because its pivot search reaches into the neighbouring block rows; this
// algorithm does the same while keeping log2(G) depth:
//
//   * pair the nodes:  z_k = (x_2k, x_2k+1)  (32 unknowns); block rows (2k+1, 2k+2) then couple
//     exactly z_k and z_k+1, i.e. the system is block BI-diagonal with 32x32 blocks
//         S_k z_k + T_k z_k+1 = f_k ,  k = 0 .. K-2,     plus 16 boundary rows at each end;
//   * one reduction level merges the row pairs (2p, 2p+1): Gaussian elimination with partial
//     pivoting over ALL 64 rows of the stacked panel [T_2p ; S_2p+1] eliminates the shared
//     unknown and leaves one reduced row (S', T') per pair (Wright's structured LU for
//     two-point boundary-value problems);  K-1 rows -> 1 row in log2 levels;
//   * the last row and the two boundary rows form a dense 64x64 system.
//
// Per merged pair the factorisation stores the row permutation, L11^-1 and L21 (forward sweep
// of a solve) and E, F, U (back substitution  z = U^-1 (g - E z_left - F z_right), done by
// substitution: an explicit U^-1 is not backward stable here).  ~66 KB per pair, ~33 KB per
// grid point and solve: the same order as LAPACK's 94 x 16 complex per column (24 KB).
//
// Solve scheduling: a CTA takes 2^mu consecutive rows of a level and reduces them to one row
// in shared memory ("stage"); rows merge pairwise, so chunks are fully independent (no halo,
// no atomics, deterministic).
#pragma once

#include <cstdint>
#include <vector>

#include "common.cuh"

namespace lgpu {

constexpr int SB = 32;                 // super-block size (two grid points)
constexpr int SB2 = SB * SB;
constexpr int TRI = SB * (SB + 1) / 2; // packed triangle
// pair record layout, in units of one complex (16 B)
constexpr int PR_PERM = 0;             // 64 x uint8
constexpr int PR_L11I = 4;             // packed lower (column-major), unit diagonal stored
constexpr int PR_L21 = PR_L11I + TRI;  // 32 x 32 column-major
constexpr int PR_E = PR_L21 + SB2;
constexpr int PR_F = PR_E + SB2;
constexpr int PR_U = PR_F + SB2;       // packed upper (column-major), diagonal stored as reciprocal
constexpr int PR_FWD_END = PR_E;
constexpr int PAIR_STRIDE = 4160;      // >= PR_U + TRI = 4132, multiple of 32
constexpr int ROW_STRIDE = 2 * SB2;    // work rows: S then T, column-major 32 x 32
// top record: perm (64 x uint8 = 4 cd), Linv 64x64, U 64x64 (reciprocal diagonal)
constexpr int TOP_LINV = 4;
constexpr int TOP_U = TOP_LINV + 64 * 64;
constexpr int TOP_STRIDE = TOP_U + 64 * 64;

struct SluLevel {
  int m;             // rows at this level
  int npairs;        // m / 2
  size_t off_pairs;  // first pair record of the level
  size_t off_rows;   // work rows (S, T) of this level
};

struct SluStage {
  int l0;            // first level of the stage
  int mu;            // levels fused (chunk = 2^mu rows of level l0)
  int m0;            // rows at level l0
  int nchunks;
  size_t off_fin;    // compact right-hand side of the stage, units of 32 complex (stage 0: b + 16)
};

struct SluPlan {
  int n = 0;          // block rows (grid points)
  int n_pad = 0;      // padded to even
  int K = 0;          // super nodes
  int top_size = 0;   // 64, or 32 when K == 1
  std::vector<SluLevel> levels;
  size_t off_rows_final = 0;   // work row holding the last reduced (S, T)
  size_t pair_records = 0;
  size_t work_rows = 0;
  std::vector<SluStage> stages;   // forward order; the last one is the single-CTA top stage
  size_t rhs_vecs = 0;
};

SluPlan make_slu_plan(int n, int first_stage_mu, int next_stage_mu, int top_max_rows);
int slu_coresident_demand(const SluPlan& plan);   // SMs a running solve may hold while waiting (see slu.cu)

// Completion flags of the last kernel of a solve (backward first stage), for a consumer that does not want to wait
// for the whole grid: chunk c of that kernel stores `epoch` into flags[c] (release) once every entry of the solution
// it owns - rows 32 * (c << mu) ... of x, boundary nodes included - is final.  `epoch` counts the solves of the
// context and never goes back; `last` is the epoch of the most recent solve IF that solve ended with such a kernel
// writing x in place (0 otherwise: the consumer must then wait for the grid).
struct SolveSignal {
  unsigned long long* flags = nullptr;   // device, >= chunks of the first stage, zero at allocation
  unsigned long long epoch = 0;          // host
  unsigned long long last = 0;           // host
  const void* x = nullptr;               // host: the vector that solve wrote
  int mu = 0;                            // rows of the first stage's chunks = 1 << mu (32 entries of x each)
  int nchunks = 0;
};

struct SluDevice {
  const cd* A;          // (n, 3, 256) blocks
  const cd* B;
  cd* pairs;            // plan.pair_records * PAIR_STRIDE
  cd* top;              // TOP_STRIDE
  cd* work;             // plan.work_rows * ROW_STRIDE
  cd* rhs;              // plan.rhs_vecs * 32
  cd* gvec;             // plan.pair_records * 32  (pivot-row right-hand sides of the last solve)
  cd* xpad;             // n_pad * 16 solution scratch when n is odd
  int32_t* info;        // singular-pivot report (1-based block row, 0 = none)
  unsigned long long* epoch;   // host: solves since the factorisation; its parity selects the mailbox
  // Mailboxes of the fused upper stages (slu_mbox_elems() complex, every byte 0xFF after a
  // factorisation): two copies (solve parity) of [stage right-hand sides | solved super nodes].
  // A CTA that needs a value another CTA of the launch produces polls the value itself until it
  // is no longer the all-ones NaN; the producer stores the value into this solve's copy and the
  // all-ones pattern into the other one, so no reader ever has to reset anything.
  cd* mbox;
  uint32_t padmask;     // bit i: row/column i of every 16-wide block belongs to a variable that is not
                        // in the state vector (hd / hd-1d): treated as a decoupled identity row
  SolveSignal* signal = nullptr;   // host; may be null
};
inline size_t slu_mbox_half(const SluPlan& p) { return (p.rhs_vecs + static_cast<size_t>(p.K)) * SB; }
inline size_t slu_mbox_elems(const SluPlan& p) { return 2 * slu_mbox_half(p); }

void slu_factorize(const SluPlan& plan, const SluDevice& d, cd sigma, cudaStream_t stream,
                   LaunchLog* log);
// x = M^-1 b ; b and x are device vectors of n*16 complex (may alias)
// With `ell` the right-hand side is b = B v, given by a compressed real-valued copy of B (bsparse.cuh) and the
// vector v: the first-stage kernel and the top system form the entries they need themselves (same order of
// operations as bell_matvec: bit-identical), and b is not read.  v must not alias x.
struct RhsEll {
  const double* val = nullptr;   // [width][rows]
  const int32_t* col = nullptr;  // [width][rows]
  const cd* x = nullptr;         // v; nullptr: plain right-hand side
  int rows = 0;
  int width = 0;                 // <= 8
};
// x_padded: x has room for n_pad * 16 entries (n odd: one padding node), the kernels write it directly and the copy
// out of the padded scratch vector - a launch between the last solve kernel and whatever follows it - is skipped
void slu_solve(const SluPlan& plan, const SluDevice& d, const cd* b, cd* x, cudaStream_t stream,
               LaunchLog* log, const RhsEll* ell = nullptr, bool x_padded = false);
// y = aa * A x + ab * B x + z : covers B*x, A*x and the refinement residual
// r = b - (A - sigma*B) x  (aa = -1, ab = sigma, z = b)
void block_matvec(int n, const cd* A, const cd* B, cd aa, cd ab, const cd* x, const cd* z, cd* y,
                  cudaStream_t stream, LaunchLog* log);

}  // namespace lgpu
