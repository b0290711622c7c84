// K1 — finite-element assembly of A and B straight into the device block-tridiagonal layout.
//
// Replaces the element loop + linked-list scatter of build_matrices
// (src/matrices/mod_matrix_manager.f08:188-260, src/matrices/mod_build_quadblock.f08:35-74,
//  src/matrices/datastructure/mod_matrix_structure.f08:65-113) and the boundary manager
// (src/boundaries/mod_boundary_manager.f08:57-84).
#pragma once

#include <cstdint>
#include <vector>

#include "../../include/legolas_b200.h"
#include "common.cuh"

namespace lgpu {

constexpr int NFIELD = LGPU_N_FIELDS;   // 38 sampled fields
constexpr int ASM_ROWS_PER_CTA = 7;     // block rows per CTA -> 8 elements = 32 Gauss points
constexpr int MASK_WORDS = 64;          // per block row: 2 matrices x 4 contributions x 256 bits
constexpr int MAX_SLOT_PER_PAIR = 4;

// One (matrix, row variable, column variable) pair and the coefficient slots feeding it.
struct PairItem {
  int32_t mat;                       // 0 = A, 1 = B
  int32_t p1, p2;                    // 0-based positions in the active state vector
  int32_t cls1, cls2;                // spline class base: 0 quadratic, 2 cubic (+ derivative flag)
  int32_t nslot;
  int32_t slot[MAX_SLOT_PER_PAIR];   // coefficient slot index
  int32_t dd[MAX_SLOT_PER_PAIR];     // d1*2 + d2 of that slot
};

// Host-built work lists for one settings combination.
struct TermPlan {
  std::vector<int32_t> slot_begin;   // nslots + 1 offsets into term_ids
  std::vector<int32_t> term_ids;     // term ids (index into terms.def)
  std::vector<PairItem> items;
  int nslots() const { return static_cast<int>(slot_begin.size()) - 1; }
};

struct AsmParams {
  int32_t gridpts;
  int32_t geometry;
  double k2, k3, gamma_1, mu, efrac;
  double nodes[4], weights[4];
};

struct DevicePlan {
  int32_t nslots, nitems;
  const int32_t* slot_begin;
  const int32_t* term_ids;
  const PairItem* items;
};

struct FieldPtrs {
  const double* f[NFIELD];
};

// Device slots (mhd positions 0..7) of the variables in the active state vector, in state-vector
// order (src/settings/mod_settings.f08:69-86); returns nb_eqs.
int state_positions(int physics_type, int slots[8]);

// Builds the element-integral plan (natural == false) or the natural-boundary plan.
TermPlan build_term_plan(const lgpu_settings& s, bool natural);
// 1-based quadblock-local indices zeroed by the essential boundary conditions
// (src/boundaries/smod_essential_boundaries.f08:12-157).
std::vector<int32_t> essential_indices(const lgpu_settings& s, bool right_edge);

void launch_assemble(const AsmParams& p, const DevicePlan& plan, const FieldPtrs& fields,
                     const double* grid, const double* gauss_grid, cd* A, cd* B, uint32_t* masks,
                     cudaStream_t stream);
void launch_boundaries(const AsmParams& p, const DevicePlan& natplan, const FieldPtrs& fields,
                       const double* grid, const double* gauss_grid, cd* A, cd* B, uint32_t* masks,
                       uint32_t* natmasks, const int32_t* ess_left, int n_left,
                       const int32_t* ess_right, int n_right, cudaStream_t stream);
size_t assemble_smem_bytes(int nslots);

}  // namespace lgpu
