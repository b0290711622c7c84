// N1 — base eigenfunctions from right eigenvectors on the device.
//
// Replaces base_ef_t%assemble (src/eigenfunctions/mod_base_efs.f08:35-61):
// assemble_eigenfunction (src/eigenfunctions/mod_ef_assembly.f08:39-106) evaluates the
// finite-element expansion of every variable on the eigenfunction grid (grid points and interval
// midpoints, src/mod_grid.f08:143-157) and retransform_eigenfunction (:16-36) undoes the scaling of
// the variables (rho, v3, T, a2: / eps; v1: / (i eps); a1: / i; v2, a3 unchanged).
#pragma once

#include "common.cuh"

namespace lgpu {

// out[(p * nsel + s) * npts + e], npts = 2 G - 1: variable p (state-vector order), selected
// eigenvector s (column idxs[s], 0-based, of the device matrix vr with leading dimension ld),
// eigenfunction-grid point e.  geometry: 0 Cartesian (eps = 1), 1 cylindrical (eps = x).
void assemble_eigenfunctions(int gridpts, int geometry, const double* grid, const cd* vr, size_t ld, int nsel,
                             const int32_t* idxs_dev, cd* out, cudaStream_t stream, LaunchLog* log);

}  // namespace lgpu
