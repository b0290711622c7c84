"""ORACLE (test infrastructure only): base eigenfunctions from right eigenvectors.

CPU restatement of
  eigenfunction grid ........... src/mod_grid.f08:143-157 (set_ef_grid)
  assemble_eigenfunction ....... src/eigenfunctions/mod_ef_assembly.f08:39-106,111-160
  retransform_eigenfunction .... src/eigenfunctions/mod_ef_assembly.f08:16-36
  base_ef_t%assemble ........... src/eigenfunctions/mod_base_efs.f08:35-61
Pinned by the reference's own stored run tests/pylbo_tests/utility_files/v2.0.0_mri_subset_efs.dat
(eigenvectors and the eigenfunctions written from them): tests/test_oracle_golden.py.
"""
import numpy as np

from . import assembly as asm

CUBIC = ("v1", "a2", "a3")


def ef_grid(grid: np.ndarray) -> np.ndarray:
    """Grid points and interval midpoints, 2 G - 1 values."""
    out = np.empty(2 * len(grid) - 1)
    out[0::2] = grid
    out[1::2] = 0.5 * (grid[:-1] + grid[1:])
    return out


def assemble_eigenfunction(name: str, state_vector, grid: np.ndarray, eigenvector: np.ndarray) -> np.ndarray:
    """Values of the finite-element expansion of variable ``name`` on the eigenfunction grid."""
    nb = 2 * len(state_vector)                       # dim_subblock
    p = list(state_vector).index(name)
    factors = asm.cubic_factors if name in CUBIC else asm.quadratic_factors
    xs = ef_grid(grid)
    out = np.empty(len(xs), dtype=np.complex128)

    def combine(idx, h):
        # idx is 0-based here; basis_function(2), (4), (1), (3) of the reference are h[1], h[3], h[0], h[2]
        return (eigenvector[idx] * h[1] + eigenvector[idx + 1] * h[3]
                + eigenvector[idx + nb] * h[0] + eigenvector[idx + nb + 1] * h[2])

    idx = 2 * p
    out[0] = combine(idx, [float(v) for v in factors(xs[0], grid[0], grid[1])])
    for g in range(len(grid) - 1):
        for e in (2 * g + 1, 2 * g + 2):             # centre and end of interval g (0-based ef index)
            out[e] = combine(idx, [float(v) for v in factors(xs[e], grid[g], grid[g + 1])])
        idx += nb
    return out


def retransform(name: str, eps: np.ndarray, ef: np.ndarray) -> np.ndarray:
    if name in ("rho", "v3", "T", "a2"):
        return ef / eps
    if name == "v1":
        return ef / (eps * 1j)
    if name in ("v2", "a3"):
        return ef.copy()
    if name == "a1":
        return ef / 1j
    raise ValueError("wrong eigenfunction name during retransform")


def base_eigenfunctions(geometry: str, state_vector, grid: np.ndarray, vr: np.ndarray, idxs):
    """{name: (2 G - 1, len(idxs)) complex}; ``idxs`` are 0-based column indices of ``vr``."""
    xs = ef_grid(grid)
    eps = xs if geometry == "cylindrical" else np.ones_like(xs)
    return {
        name: np.stack([retransform(name, eps, assemble_eigenfunction(name, state_vector, grid, vr[:, k]))
                        for k in idxs], axis=1)
        for name in state_vector
    }
