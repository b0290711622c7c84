"""NumPy restatement of Legolas' finite-element assembly (``build_matrices``).

Test infrastructure (see ``oracle/__init__.py``): the checker for the CUDA
assembly kernel, never the thing shipped.  Vectorised over grid intervals but
keeps the reference's *per-element summation order*: Gauss point -> physics
procedure -> term, ``spline1 * (weight*factor) * spline2`` left to right, then
the ``dx`` scaling, then the per-contribution drop rule on insertion.

Follows (reference file:line):
  element loop ........... src/matrices/mod_matrix_manager.f08:138-266
  term expansion ......... src/matrices/mod_build_quadblock.f08:19-78
  skip-absent-variable ... src/matrices/elements/mod_matrix_elements.f08:46-81
  splines ................ src/mod_spline_functions.f08:23-99
  Gauss constants ........ src/mod_global_variables.f08:29-43
  grids .................. src/mod_grid.f08:121-140,160-227
  B / regular A terms .... src/matrices/smod_regular_matrix.f08:6-163
  flow ................... src/matrices/smod_flow_matrix.f08:6-123
  resistive .............. src/matrices/smod_resistive_matrix.f08:6-165
  heat-loss .............. src/matrices/smod_heatloss_matrix.f08:6-27
  conduction ............. src/matrices/smod_conduction_matrix.f08:6-263
  viscosity .............. src/matrices/smod_viscosity_matrix.f08:6-173
  Hall ................... src/matrices/smod_hall_matrix.f08:6-368
  boundary manager ....... src/boundaries/mod_boundary_manager.f08:57-107
  natural boundaries ..... src/boundaries/smod_natural_boundaries.f08:103-222,
                           src/boundaries/smod_natural_bounds_*.f08
  essential boundaries ... src/boundaries/smod_essential_boundaries.f08:12-205
  drop rule (is_zero) .... src/matrices/datastructure/mod_matrix_structure.f08:65-113,
                           src/mod_check_values.f08:143-192
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

DP_LIMIT = 5.0e-15  # src/mod_global_variables.f08:19
IC = 1j

# src/mod_global_variables.f08:31-43 (Legolas 2.0.6, proper double literals)
GAUSS_NODES = np.array(
    [-0.861136311594053, -0.339981043584856, 0.339981043584856, 0.861136311594053]
)
GAUSS_WEIGHTS = np.array(
    [0.347854845137454, 0.652145154862546, 0.652145154862546, 0.347854845137454]
)
# Every stored golden datfile (<= 2.0.0) was produced with single-precision
# literals: nodes/weights = float64(float32(value)).
LEGACY_GAUSS_NODES = GAUSS_NODES.astype(np.float32).astype(np.float64)
LEGACY_GAUSS_WEIGHTS = GAUSS_WEIGHTS.astype(np.float32).astype(np.float64)

STATE_VECTORS = {
    "mhd": ("rho", "v1", "v2", "v3", "T", "a1", "a2", "a3"),
    "hd": ("rho", "v1", "v2", "v3", "T"),
    "hd-1d": ("rho", "v1", "T"),
}

# Field slots sampled at the Gaussian grid; this is also the C-ABI slot order
# (include/legolas_b200.h).  Absent == identically zero.
FIELD_NAMES = (
    "rho0", "drho0", "T0", "dT0", "ddT0",
    "B01", "B02", "dB02", "ddB02", "B03", "dB03", "ddB03",
    "v01", "dv01", "ddv01", "v02", "dv02", "ddv02", "v03", "dv03", "ddv03",
    "g0", "eta", "detadT", "detadr", "L0", "dLdT", "dLdrho",
    "tcpara", "dtcparadT", "tcperp", "dtcperpdrho", "dtcperpdT", "dtcperpdB2",
    "tcprefactor", "dtcprefactordr", "hallfactor", "inertiafactor",
)


@dataclass
class Settings:
    """The scalars of ``settings_t`` / ``mod_equilibrium_params`` the assembly reads."""

    gridpts: int
    geometry: str = "Cartesian"          # "Cartesian" | "cylindrical"
    physics_type: str = "mhd"
    k2: float = 0.0
    k3: float = 0.0
    gamma: float = 5.0 / 3.0
    incompressible: bool = False
    flow: bool = False
    resistivity: bool = False
    cooling: bool = False
    heating: bool = False
    conduction: bool = False             # any conduction (parallel or perpendicular)
    perpendicular_conduction: bool = False
    viscosity: bool = False
    viscosity_value: float = 0.0
    viscous_heating: bool = False
    hall: bool = False
    electron_inertia: bool = False
    electron_fraction: float = 0.5
    gravity: bool = False
    boundary_type: str = "wall"          # "wall" | "wall_weak"
    coaxial: bool = False
    gauss_nodes: np.ndarray = field(default_factory=lambda: GAUSS_NODES.copy())
    gauss_weights: np.ndarray = field(default_factory=lambda: GAUSS_WEIGHTS.copy())

    @property
    def state_vector(self):
        return STATE_VECTORS[self.physics_type]

    @property
    def nb_eqs(self):
        return len(self.state_vector)

    @property
    def dim_subblock(self):          # src/settings/mod_dims.f08:36-44
        return 2 * self.nb_eqs

    @property
    def dim_quadblock(self):
        return 4 * self.nb_eqs

    @property
    def dim_matrix(self):
        return self.gridpts * self.dim_subblock

    @property
    def gamma_1(self):
        # incompressible sets gamma = 1e12 (src/settings/mod_physics_settings.f08:90-94)
        return (1.0e12 if self.incompressible else self.gamma) - 1.0

    @property
    def has_bfield(self):
        return self.physics_type == "mhd"


# --------------------------------------------------------------------------- grids
def base_grid(start: float, end: float, gridpts: int) -> np.ndarray:
    """Uniform base grid with the reference's accumulation (src/mod_grid.f08:160-196)."""
    dx = (end - start) / (gridpts - 1)
    xbar = np.empty(gridpts)
    xbar[0] = start
    for i in range(1, gridpts):
        xbar[i] = xbar[i - 1] + dx
    kappa = (end - xbar[gridpts - 2]) / (xbar[gridpts - 1] - xbar[gridpts - 2])
    grid = np.empty(gridpts)
    grid[0] = start
    grid[1:] = xbar[:-1] + kappa * dx
    return grid


def gaussian_grid(grid: np.ndarray, nodes: np.ndarray = GAUSS_NODES) -> np.ndarray:
    """src/mod_grid.f08:121-140."""
    x_lo, x_hi = grid[:-1], grid[1:]
    dx = x_hi - x_lo
    out = 0.5 * dx[:, None] * nodes[None, :] + 0.5 * (x_lo + x_hi)[:, None]
    return out.reshape(-1)


# ------------------------------------------------------------------------- splines
def quadratic_factors(r, lo, hi):
    """src/mod_spline_functions.f08:23-37."""
    z = np.zeros_like(r + lo)
    return [
        4.0 * (r - lo) * (hi - r) / (hi - lo) ** 2,
        z,
        (2.0 * r - hi - lo) * (r - lo) / (hi - lo) ** 2,
        (2.0 * r - hi - lo) * (r - hi) / (hi - lo) ** 2,
    ]


def quadratic_factors_deriv(r, lo, hi):
    """src/mod_spline_functions.f08:41-56."""
    z = np.zeros_like(r + lo)
    return [
        4.0 * (-2.0 * r + hi + lo) / (hi - lo) ** 2,
        z,
        (4.0 * r - hi - 3.0 * lo) / (hi - lo) ** 2,
        (4.0 * r - lo - 3.0 * hi) / (hi - lo) ** 2,
    ]


def cubic_factors(r, lo, hi):
    """src/mod_spline_functions.f08:60-76."""
    return [
        3.0 * ((r - lo) / (hi - lo)) ** 2 - 2.0 * ((r - lo) / (hi - lo)) ** 3,
        3.0 * ((hi - r) / (hi - lo)) ** 2 - 2.0 * ((hi - r) / (hi - lo)) ** 3,
        (r - hi) * ((r - lo) / (hi - lo)) ** 2,
        (r - lo) * ((hi - r) / (hi - lo)) ** 2,
    ]


def cubic_factors_deriv(r, lo, hi):
    """src/mod_spline_functions.f08:80-99."""
    return [
        6.0 * (r - lo) / (hi - lo) ** 2 - 6.0 * (r - lo) ** 2 / (hi - lo) ** 3,
        -6.0 * (hi - r) / (hi - lo) ** 2 + 6.0 * (hi - r) ** 2 / (hi - lo) ** 3,
        (2.0 * (r - hi) * (r - lo) + (r - lo) ** 2) / (hi - lo) ** 2,
        (2.0 * (r - lo) * (r - hi) + (r - hi) ** 2) / (hi - lo) ** 2,
    ]


def _splines(r, lo, hi):
    return {
        "h_quad": quadratic_factors(r, lo, hi),
        "dh_quad": quadratic_factors_deriv(r, lo, hi),
        "h_cubic": cubic_factors(r, lo, hi),
        "dh_cubic": cubic_factors_deriv(r, lo, hi),
    }


# ------------------------------------------------------------------- term collector
class Elements:
    """Term list of one procedure at one Gauss point (matrix_elements_t)."""

    def __init__(self, state_vector):
        self.state_vector = state_vector
        self.terms = []

    def add(self, factor, loc1, loc2, spline1, spline2):
        # absent variable -> term silently skipped (mod_matrix_elements.f08:57-59)
        if loc1 not in self.state_vector or loc2 not in self.state_vector:
            return
        p1 = self.state_vector.index(loc1) + 1
        p2 = self.state_vector.index(loc2) + 1
        self.terms.append((factor, p1, p2, spline1, spline2))


def add_to_quadblock(quadblock, elements, weight, dim_subblock, spl):
    """mod_build_quadblock.f08:35-74.  quadblock has shape (dimq, dimq, n)."""
    # spline entry -> row offset: entries (2,4) -> left node rows (2p-1, 2p),
    # entries (1,3) -> right node rows (+dim_subblock)
    order = (1, 3, 0, 2)
    for factor, p1, p2, s1name, s2name in elements.terms:
        s1, s2 = spl[s1name], spl[s2name]
        fac = weight * (np.asarray(factor) + 0j)
        rows = (2 * p1 - 2, 2 * p1 - 1, 2 * p1 - 2 + dim_subblock, 2 * p1 - 1 + dim_subblock)
        cols = (2 * p2 - 2, 2 * p2 - 1, 2 * p2 - 2 + dim_subblock, 2 * p2 - 1 + dim_subblock)
        for a in range(4):
            left = s1[order[a]] * fac
            for b in range(4):
                quadblock[rows[a], cols[b]] += left * s2[order[b]]


def _eps(s: Settings, x):
    if s.geometry == "Cartesian":
        return np.ones_like(x), 0.0
    return x, 1.0


# ------------------------------------------------------------------ physics terms
def add_bmatrix_terms(x, s, f):
    """smod_regular_matrix.f08:6-28."""
    rho = f["rho0"]
    eps, _ = _eps(s, x)
    one = np.ones_like(x)
    el = Elements(s.state_vector)
    el.add(one, "rho", "rho", "h_quad", "h_quad")
    el.add(eps * rho, "v2", "v2", "h_quad", "h_quad")
    el.add(rho, "v3", "v3", "h_quad", "h_quad")
    el.add(rho, "T", "T", "h_quad", "h_quad")
    el.add(eps, "a1", "a1", "h_quad", "h_quad")
    el.add(rho, "v1", "v1", "h_cubic", "h_cubic")
    el.add(one, "a2", "a2", "h_cubic", "h_cubic")
    el.add(eps, "a3", "a3", "h_cubic", "h_cubic")
    return el


def add_regular_matrix_terms(x, s, f):
    """smod_regular_matrix.f08:31-163."""
    k2, k3 = s.k2, s.k3
    gamma_1 = s.gamma_1
    eps, deps = _eps(s, x)
    rho, drho = f["rho0"], f["drho0"]
    T0, dT0 = f["T0"], f["dT0"]
    g0 = f["g0"]
    B01, B02, dB02, B03, dB03 = f["B01"], f["B02"], f["dB02"], f["B03"], f["dB03"]
    drB02 = deps * B02 + eps * dB02
    Fop_plus = k2 * B02 / eps + k3 * B03
    Gop_plus = k3 * B02 + k2 * B03 / eps
    Gop_min = k3 * B02 - k2 * B03 / eps
    WVop = k2**2 / eps + eps * k3**2

    el = Elements(s.state_vector)
    # Quadratic * Cubic
    el.add(-drho, "rho", "v1", "h_quad", "h_cubic")
    el.add(k3 * (drB02 - IC * k2 * B01) / eps, "v2", "a2", "h_quad", "h_cubic")
    el.add(k2 * (IC * k2 * B01 - drB02) / eps, "v2", "a3", "h_quad", "h_cubic")
    el.add(k3 * (dB03 - IC * k3 * B01), "v3", "a2", "h_quad", "h_cubic")
    el.add(k2 * (IC * k3 * B01 - dB03), "v3", "a3", "h_quad", "h_cubic")
    if not s.incompressible:
        el.add(-dT0 * rho, "T", "v1", "h_quad", "h_cubic")
    # Quadratic * dCubic
    el.add(-rho, "rho", "v1", "h_quad", "dh_cubic")
    el.add(k2 * B03 / eps, "v2", "a2", "h_quad", "dh_cubic")
    el.add(eps * k3 * B03, "v2", "a3", "h_quad", "dh_cubic")
    el.add(-(k2 * B02 + IC * deps * B01) / eps, "v3", "a2", "h_quad", "dh_cubic")
    el.add(-eps * k3 * B02, "v3", "a3", "h_quad", "dh_cubic")
    el.add(-gamma_1 * T0 * rho, "T", "v1", "h_quad", "dh_cubic")
    # Quadratic * Quadratic
    el.add(rho * k2, "rho", "v2", "h_quad", "h_quad")
    el.add(rho * k3, "rho", "v3", "h_quad", "h_quad")
    el.add(k2 * T0 / eps, "v2", "rho", "h_quad", "h_quad")
    el.add(k2 * rho / eps, "v2", "T", "h_quad", "h_quad")
    el.add(-WVop * B03, "v2", "a1", "h_quad", "h_quad")
    el.add(k3 * T0, "v3", "rho", "h_quad", "h_quad")
    el.add(k3 * rho, "v3", "T", "h_quad", "h_quad")
    el.add(IC * deps * k2 * B01 / eps + B02 * WVop, "v3", "a1", "h_quad", "h_quad")
    el.add(gamma_1 * k2 * rho * T0, "T", "v2", "h_quad", "h_quad")
    el.add(gamma_1 * k3 * rho * T0, "T", "v3", "h_quad", "h_quad")
    el.add(-eps * B03, "a1", "v2", "h_quad", "h_quad")
    el.add(B02, "a1", "v3", "h_quad", "h_quad")
    # Cubic * Quadratic
    el.add(-deps * T0 / eps, "v1", "rho", "h_cubic", "h_quad")
    if s.gravity:
        el.add(g0, "v1", "rho", "h_cubic", "h_quad")
    el.add(-deps * rho / eps, "v1", "T", "h_cubic", "h_quad")
    el.add(deps * Gop_plus, "v1", "a1", "h_cubic", "h_quad")
    el.add(IC * B01, "a2", "v3", "h_cubic", "h_quad")
    el.add(-IC * eps * B01, "a3", "v2", "h_cubic", "h_quad")
    # dCubic * Quadratic
    el.add(-T0, "v1", "rho", "dh_cubic", "h_quad")
    el.add(-rho, "v1", "T", "dh_cubic", "h_quad")
    el.add(-eps * Gop_min, "v1", "a1", "dh_cubic", "h_quad")
    # Cubic * Cubic
    el.add(-k3 * Fop_plus, "v1", "a2", "h_cubic", "h_cubic")
    el.add(k2 * Fop_plus, "v1", "a3", "h_cubic", "h_cubic")
    el.add(-B03, "a2", "v1", "h_cubic", "h_cubic")
    el.add(B02, "a3", "v1", "h_cubic", "h_cubic")
    # Cubic * dCubic
    el.add(-deps * B03 / eps, "v1", "a2", "h_cubic", "dh_cubic")
    el.add(-deps * B02, "v1", "a3", "h_cubic", "dh_cubic")
    # dCubic * dCubic
    el.add(-B03, "v1", "a2", "dh_cubic", "dh_cubic")
    el.add(eps * B02, "v1", "a3", "dh_cubic", "dh_cubic")
    # dQuadratic * Quadratic
    el.add(-IC * eps * k3 * B01, "v2", "a1", "dh_quad", "h_quad")
    el.add(IC * k2 * B01, "v3", "a1", "dh_quad", "h_quad")
    # dQuadratic * dCubic
    el.add(IC * eps * B01, "v2", "a3", "dh_quad", "dh_cubic")
    el.add(-IC * B01, "v3", "a2", "dh_quad", "dh_cubic")
    return el


def add_flow_matrix_terms(x, s, f):
    """smod_flow_matrix.f08:6-123."""
    k2, k3 = s.k2, s.k3
    gamma_1 = s.gamma_1
    eps, deps = _eps(s, x)
    rho, drho, T0 = f["rho0"], f["drho0"], f["T0"]
    v01, dv01 = f["v01"], f["dv01"]
    drv01 = deps * v01 + eps * dv01
    v02, dv02 = f["v02"], f["dv02"]
    drv02 = deps * v02 + eps * dv02
    v03, dv03 = f["v03"], f["dv03"]
    Vop = k2 * v02 / eps + k3 * v03

    el = Elements(s.state_vector)
    # Quadratic * Quadratic
    el.add(Vop - IC * dv01, "rho", "rho", "h_quad", "h_quad")
    el.add(-drv02 * IC * v01 / eps, "v2", "rho", "h_quad", "h_quad")
    el.add(rho * (eps * Vop - IC * deps * v01), "v2", "v2", "h_quad", "h_quad")
    el.add(-IC * v01 * dv03, "v3", "rho", "h_quad", "h_quad")
    el.add(rho * (Vop + IC * dv01) + (deps * rho / eps + drho) * IC * v01,
           "v3", "v3", "h_quad", "h_quad")
    el.add(eps * Vop, "a1", "a1", "h_quad", "h_quad")
    # Quadratic * dQuadratic
    el.add(-IC * v01, "rho", "rho", "h_quad", "dh_quad")
    el.add(-IC * eps * rho * v01, "v2", "v2", "h_quad", "dh_quad")
    # Cubic * Quadratic
    el.add(v01 * dv01 - deps * v02**2 / eps, "v1", "rho", "h_cubic", "h_quad")
    el.add(-2.0 * deps * rho * v02, "v1", "v2", "h_cubic", "h_quad")
    el.add(IC * v01 * k2, "a2", "a1", "h_cubic", "h_quad")
    el.add(eps * k3 * IC * v01, "a3", "a1", "h_cubic", "h_quad")
    # Cubic * Cubic
    el.add(rho * Vop + (deps * rho / eps + drho) * IC * v01, "v1", "v1", "h_cubic", "h_cubic")
    el.add(k3 * v03, "a2", "a2", "h_cubic", "h_cubic")
    el.add(-k2 * v03, "a2", "a3", "h_cubic", "h_cubic")
    el.add(-k3 * v02, "a3", "a2", "h_cubic", "h_cubic")
    el.add(k2 * v02, "a3", "a3", "h_cubic", "h_cubic")
    # dCubic * Cubic
    el.add(IC * rho * v01, "v1", "v1", "dh_cubic", "h_cubic")
    # Quadratic * Cubic
    el.add(-drv02 * rho / eps, "v2", "v1", "h_quad", "h_cubic")
    el.add(-rho * dv03, "v3", "v1", "h_quad", "h_cubic")
    # dQuadratic * Quadratic
    el.add(IC * rho * v01, "v3", "v3", "dh_quad", "h_quad")
    # Quadratic * dCubic
    el.add(-v02, "a1", "a2", "h_quad", "dh_cubic")
    el.add(-eps * v03, "a1", "a3", "h_quad", "dh_cubic")
    # Cubic * dCubic
    el.add(-IC * v01, "a2", "a2", "h_cubic", "dh_cubic")
    el.add(-IC * v01, "a3", "a3", "h_cubic", "dh_cubic")
    if not s.incompressible:
        el.add(-IC * gamma_1 * drv01 * T0 / eps, "T", "rho", "h_quad", "h_quad")
        el.add(rho * (Vop + IC * dv01 - IC * gamma_1 * drv01 / eps)
               + IC * v01 * (deps * rho / eps + drho), "T", "T", "h_quad", "h_quad")
        el.add(IC * rho * v01, "T", "T", "dh_quad", "h_quad")
    return el


def add_resistive_matrix_terms(x, s, f):
    """smod_resistive_matrix.f08:6-165."""
    k2, k3 = s.k2, s.k3
    gamma_1 = s.gamma_1
    eps, deps = _eps(s, x)
    B02, dB02, ddB02 = f["B02"], f["dB02"], f["ddB02"]
    drB02 = deps * B02 + eps * dB02
    B03, dB03, ddB03 = f["B03"], f["dB03"], f["ddB03"]
    eta, detadT = f["eta"], f["detadT"]
    deta = f["detadr"] + (f["dT0"] * detadT)
    WVop = k2**2 / eps + eps * k3**2
    Rop_pos = deps * eta / eps + deta
    Rop_neg = deps * eta / eps - deta

    el = Elements(s.state_vector)
    el.add(-IC * eta * WVop, "a1", "a1", "h_quad", "h_quad")
    el.add(IC * eta * k2 / eps, "a1", "a2", "h_quad", "dh_cubic")
    el.add(IC * eta * eps * k3, "a1", "a3", "h_quad", "dh_cubic")
    el.add(IC * dB03 * detadT, "a2", "T", "h_cubic", "h_quad")
    el.add(IC * k2 * Rop_pos, "a2", "a1", "h_cubic", "h_quad")
    el.add(-IC * drB02 * detadT / eps, "a3", "T", "h_cubic", "h_quad")
    el.add(IC * deta * eps * k3, "a3", "a1", "h_cubic", "h_quad")
    el.add(IC * eta * k2, "a2", "a1", "dh_cubic", "h_quad")
    el.add(IC * eta * eps * k3, "a3", "a1", "dh_cubic", "h_quad")
    el.add(-IC * eta * k3**2, "a2", "a2", "h_cubic", "h_cubic")
    el.add(IC * eta * k2 * k3, "a2", "a3", "h_cubic", "h_cubic")
    el.add(IC * eta * k2 * k3 / eps, "a3", "a2", "h_cubic", "h_cubic")
    el.add(-IC * eta * k2**2 / eps, "a3", "a3", "h_cubic", "h_cubic")
    el.add(-IC * Rop_pos, "a2", "a2", "h_cubic", "dh_cubic")
    el.add(-IC * deta * eps, "a3", "a3", "h_cubic", "dh_cubic")
    el.add(-IC * eta, "a2", "a2", "dh_cubic", "dh_cubic")
    el.add(-IC * eta * eps, "a3", "a3", "dh_cubic", "dh_cubic")
    if not s.incompressible:
        el.add(IC * gamma_1 * detadT * ((drB02 / eps) ** 2 + dB03**2),
               "T", "T", "h_quad", "h_quad")
        el.add(2.0 * IC * gamma_1 * (
            k2 * (dB03 * Rop_pos + eta * ddB03)
            + k3 * (drB02 * Rop_neg - eta * (2.0 * deps * dB02 + eps * ddB02))
        ), "T", "a1", "h_quad", "h_quad")
        el.add(-2.0 * IC * gamma_1 * eta * (k3 * drB02 - k2 * dB03),
               "T", "a1", "dh_quad", "h_quad")
        el.add(-2.0 * IC * gamma_1 * eta * (drB02 * k2 * k3 / eps**2 + k3**2 * dB03),
               "T", "a2", "h_quad", "h_cubic")
        el.add(2.0 * IC * gamma_1 * eta * (drB02 * k2**2 / eps**2 + k2 * k3 * dB03),
               "T", "a3", "h_quad", "h_cubic")
        el.add(-2.0 * IC * gamma_1 * (dB03 * Rop_pos + ddB03 * eta),
               "T", "a2", "h_quad", "dh_cubic")
        el.add(-2.0 * IC * gamma_1 * (
            drB02 * Rop_neg - eta * (2.0 * deps * dB02 + eps * ddB02)
        ), "T", "a3", "h_quad", "dh_cubic")
        el.add(-2.0 * IC * gamma_1 * eta * dB03, "T", "a2", "dh_quad", "dh_cubic")
        el.add(2.0 * IC * gamma_1 * drB02 * eta, "T", "a3", "dh_quad", "dh_cubic")
    return el


def add_heatloss_matrix_terms(x, s, f):
    """smod_heatloss_matrix.f08:6-27."""
    el = Elements(s.state_vector)
    if s.incompressible:
        return el
    gamma_1 = s.gamma_1
    rho = f["rho0"]
    Lrho, LT, L0 = f["dLdrho"], f["dLdT"], f["L0"]
    el.add(-IC * gamma_1 * (L0 + rho * Lrho), "T", "rho", "h_quad", "h_quad")
    el.add(-IC * gamma_1 * rho * LT, "T", "T", "h_quad", "h_quad")
    return el


def _B0(f):
    # src/background/fields/mod_bg_magnetic.f08:38-42
    return np.sqrt(f["B01"] ** 2 + f["B02"] ** 2 + f["B03"] ** 2)


def add_conduction_matrix_terms(x, s, f):
    """smod_conduction_matrix.f08:6-263."""
    el = Elements(s.state_vector)
    if s.incompressible:
        return el
    k2, k3 = s.k2, s.k3
    gamma_1 = s.gamma_1
    eps, deps = _eps(s, x)
    dT0 = f["dT0"]
    kappa_perp = f["tcperp"]
    dkappa_perp_drho = f["dtcperpdrho"]
    dkappa_perp_dT = f["dtcperpdT"]
    WVop = k2**2 / eps + eps * k3**2

    el.add(-gamma_1 * IC * WVop * kappa_perp / eps, "T", "T", "h_quad", "h_quad")
    el.add(-IC * gamma_1 * dT0 * dkappa_perp_drho, "T", "rho", "dh_quad", "h_quad")
    el.add(gamma_1 * (IC * deps * kappa_perp / eps - IC * dT0 * dkappa_perp_dT),
           "T", "T", "dh_quad", "h_quad")
    el.add(-IC * gamma_1 * kappa_perp, "T", "T", "dh_quad", "dh_quad")
    if not s.has_bfield:
        return el

    ddT0 = f["ddT0"]
    B0 = _B0(f)
    B01, B02, dB02, B03, dB03 = f["B01"], f["B02"], f["dB02"], f["B03"], f["dB03"]
    dkappa_para_dT = f["dtcparadT"]
    dkappa_perp_dB2 = f["dtcperpdB2"]
    Kp = f["tcprefactor"]
    diffKp = f["dtcprefactordr"]
    Kp_plus = Kp + dkappa_perp_dB2
    Kp_plusplus = dkappa_perp_dB2 - (B01**2 * Kp_plus / B0**2)
    Fop_plus = k2 * B02 / eps + k3 * B03
    dFop_plus = (k2 / eps) * (dB02 - deps * B02 / eps) + k3 * dB03
    Gop_min = k3 * B02 - k2 * B03 / eps
    Fop_B01 = deps * IC * B01 / eps + Fop_plus

    el.add(gamma_1 * dT0 * (B01 / B0**2) * dkappa_perp_drho * Fop_B01,
           "T", "rho", "h_quad", "h_quad")
    el.add(gamma_1 * (
        B01 * Fop_plus * diffKp
        - B01 * dT0 * (dkappa_para_dT - dkappa_perp_dT) * Fop_B01 / B0**2
        + Kp * (
            B01 * deps * (2.0 * deps * IC * B01 / eps + 3.0 * Fop_plus) / eps
            + B01 * dFop_plus
            - IC * Fop_plus**2
        )
    ), "T", "T", "h_quad", "h_quad")
    el.add(2.0 * gamma_1 * eps * dT0 * Gop_min * Kp_plus * B01 * Fop_B01 / B0**2,
           "T", "a1", "h_quad", "h_quad")
    el.add(IC * gamma_1 * dT0 * dkappa_perp_drho * B01**2 / B0**2,
           "T", "rho", "dh_quad", "h_quad")
    el.add(gamma_1 * (
        B01 * Kp * (2.0 * deps * IC * B01 / eps + 3.0 * Fop_plus)
        - IC * dT0 * (B01**2 * dkappa_para_dT / B0**2 - dkappa_perp_dT * B01**2 / B0**2)
    ), "T", "T", "dh_quad", "h_quad")
    el.add(-2.0 * IC * gamma_1 * eps * dT0 * Gop_min * Kp_plusplus,
           "T", "a1", "dh_quad", "h_quad")
    el.add(-2.0 * IC * gamma_1 * deps * B01**2 * Kp / eps, "T", "T", "h_quad", "dh_quad")
    el.add(-IC * gamma_1 * 2.0 * B01**2 * Kp, "T", "T", "dh_quad", "dh_quad")
    common = (
        Kp * (B01 * ddT0 + IC * dT0 * Fop_plus)
        + B01 * dT0 * (diffKp - 2.0 * IC * B01 * Kp_plus * Fop_B01 / B0**2)
    )
    el.add(gamma_1 * k3 * common, "T", "a2", "h_quad", "h_cubic")
    el.add(-gamma_1 * k2 * common, "T", "a3", "h_quad", "h_cubic")
    el.add(gamma_1 * dT0 * B01 * (2.0 * B03 * Kp_plus * Fop_B01 / B0**2 - k3 * Kp),
           "T", "a2", "h_quad", "dh_cubic")
    el.add(gamma_1 * dT0 * B01 * (-2.0 * eps * B02 * Kp_plus * Fop_B01 / B0**2 + k2 * Kp),
           "T", "a3", "h_quad", "dh_cubic")
    el.add(-2.0 * gamma_1 * dT0 * B01 * k3 * Kp_plusplus, "T", "a2", "dh_quad", "h_cubic")
    el.add(2.0 * gamma_1 * dT0 * B01 * k2 * Kp_plusplus, "T", "a3", "dh_quad", "h_cubic")
    el.add(-2.0 * IC * gamma_1 * dT0 * B03 * Kp_plusplus, "T", "a2", "dh_quad", "dh_cubic")
    el.add(2.0 * IC * gamma_1 * dT0 * eps * B02 * Kp_plusplus, "T", "a3", "dh_quad", "dh_cubic")
    return el


def add_viscosity_matrix_terms(x, s, f):
    """smod_viscosity_matrix.f08:6-173."""
    k2, k3 = s.k2, s.k3
    gamma_1 = s.gamma_1
    eps, deps = _eps(s, x)
    v01, dv01, ddv01 = f["v01"], f["dv01"], f["ddv01"]
    v02, dv02 = f["v02"], f["dv02"]
    dv03, ddv03 = f["dv03"], f["ddv03"]
    mu = s.viscosity_value
    WVop = k2**2 / eps + eps * k3**2
    one = np.ones_like(x)

    el = Elements(s.state_vector)
    el.add(-IC * mu * (deps / eps + WVop) / eps, "v1", "v1", "h_cubic", "h_cubic")
    el.add(-IC * mu * deps / (3.0 * eps), "v1", "v1", "h_cubic", "dh_cubic")
    el.add(IC * mu * deps / eps, "v1", "v1", "dh_cubic", "h_cubic")
    el.add(-4.0 * IC * mu / 3.0 * one, "v1", "v1", "dh_cubic", "dh_cubic")
    el.add(7.0 * deps * IC * mu * k2 / (3.0 * eps), "v1", "v2", "h_cubic", "h_quad")
    el.add(IC * mu * deps * k3 / (3.0 * eps), "v1", "v3", "h_cubic", "h_quad")
    el.add(IC * mu * k2 / 3.0 * one, "v1", "v2", "dh_cubic", "h_quad")
    el.add(IC * mu * k3 / 3.0 * one, "v1", "v3", "dh_cubic", "h_quad")
    el.add(IC * mu * deps * 2.0 * k2 / eps**2, "v2", "v1", "h_quad", "h_cubic")
    el.add(IC * mu * k2 / (3.0 * eps), "v2", "v1", "h_quad", "dh_cubic")
    el.add(IC * mu * k3 / 3.0 * one, "v3", "v1", "h_quad", "dh_cubic")
    el.add(-IC * mu * (deps / eps + 4.0 * k2**2 / (3.0 * eps) + eps * k3**2),
           "v2", "v2", "h_quad", "h_quad")
    el.add(-IC * mu * k2 * k3 / (3.0 * eps), "v2", "v3", "h_quad", "h_quad")
    el.add(-IC * mu * k2 * k3 / 3.0 * one, "v3", "v2", "h_quad", "h_quad")
    el.add(-IC * mu * (k2**2 / eps**2 + 4.0 * k3**2 / 3.0), "v3", "v3", "h_quad", "h_quad")
    el.add(-IC * mu * eps, "v2", "v2", "dh_quad", "dh_quad")
    el.add(-IC * mu * one, "v3", "v3", "dh_quad", "dh_quad")
    el.add(IC * mu * deps / eps, "v3", "v3", "dh_quad", "h_quad")
    if s.viscous_heating and not s.incompressible:
        el.add(2.0 * gamma_1 * mu * (
            (deps**2 * v01 - IC * deps * k2 * v02) / eps**2 - deps * dv01 / eps - ddv01
        ), "T", "v1", "h_quad", "h_cubic")
        el.add(2.0 * gamma_1 * mu * (deps**2 * IC * v02 - deps * k2 * v01) / eps,
               "T", "v2", "h_quad", "h_quad")
        el.add(-2.0 * IC * gamma_1 * mu * (deps * dv03 / eps + ddv03),
               "T", "v3", "h_quad", "h_quad")
        el.add(-2.0 * IC * gamma_1 * mu * dv03, "T", "v3", "dh_quad", "h_quad")
        el.add(-2.0 * gamma_1 * mu * dv01, "T", "v1", "dh_quad", "h_cubic")
        el.add(2.0 * IC * gamma_1 * mu * eps * dv02, "T", "v2", "h_quad", "dh_quad")
    return el


def add_hall_bmatrix_terms(x, s, f):
    """smod_hall_matrix.f08:6-86."""
    k2, k3 = s.k2, s.k3
    eps, deps = _eps(s, x)
    rho, drho = f["rho0"], f["drho0"]
    eta_H, eta_e = f["hallfactor"], f["inertiafactor"]
    WVop = k2**2 / eps + eps * k3**2
    el = Elements(s.state_vector)
    el.add(eta_H, "a1", "v1", "h_quad", "h_cubic")
    el.add(eta_H * eps, "a2", "v2", "h_cubic", "h_quad")
    el.add(eta_H, "a3", "v3", "h_cubic", "h_quad")
    if s.electron_inertia:
        el.add(eta_e * WVop / rho, "a1", "a1", "h_quad", "h_quad")
        el.add(-eta_e * k2 / (eps * rho), "a1", "a2", "h_quad", "dh_cubic")
        el.add(-eta_e * eps * k3 / rho, "a1", "a3", "h_quad", "dh_cubic")
        el.add(-eta_e * k2 * (deps / (eps * rho) - drho / rho**2),
               "a2", "a1", "h_cubic", "h_quad")
        el.add(eta_e * drho * eps * k3 / rho**2, "a3", "a1", "h_cubic", "h_quad")
        el.add(-eta_e * k2 / rho, "a2", "a1", "dh_cubic", "h_quad")
        el.add(-eta_e * eps * k3 / rho, "a3", "a1", "dh_cubic", "h_quad")
        el.add(eta_e * k3**2 / rho, "a2", "a2", "h_cubic", "h_cubic")
        el.add(-eta_e * k2 * k3 / rho, "a2", "a3", "h_cubic", "h_cubic")
        el.add(-eta_e * k2 * k3 / (eps * rho), "a3", "a2", "h_cubic", "h_cubic")
        el.add(eta_e * k2**2 / (eps * rho), "a3", "a3", "h_cubic", "h_cubic")
        el.add(eta_e * (deps / (eps * rho) - drho / rho**2), "a2", "a2", "h_cubic", "dh_cubic")
        el.add(-eta_e * eps * drho / rho**2, "a3", "a3", "h_cubic", "dh_cubic")
        el.add(eta_e / rho, "a2", "a2", "dh_cubic", "dh_cubic")
        el.add(eta_e * eps / rho, "a3", "a3", "dh_cubic", "dh_cubic")
    return el


def add_hall_matrix_terms(x, s, f):
    """smod_hall_matrix.f08:89-368.

    Quirk replicated: without viscosity the procedure returns *before*
    ``add_to_quadblock`` (:191), so nothing at all is added.
    """
    el = Elements(s.state_vector)
    if not s.viscosity:
        return el
    k2, k3 = s.k2, s.k3
    eps, deps = _eps(s, x)
    v01, v02, v03 = f["v01"], f["v02"], f["v03"]
    dv01, dv02, dv03 = f["dv01"], f["dv02"], f["dv03"]
    ddv01, ddv02, ddv03 = f["ddv01"], f["ddv02"], f["ddv03"]
    rho, drho, dT0 = f["rho0"], f["drho0"], f["dT0"]
    eta_H = f["hallfactor"]
    mu = s.viscosity_value
    efrac = s.electron_fraction

    el.add(eta_H * (k2 * v02 / eps + k3 * v03), "a1", "v1", "h_quad", "h_cubic")
    el.add(-eta_H * (1.0 - efrac) * dT0 / rho, "a1", "rho", "h_quad", "h_quad")
    el.add(-2.0 * eta_H * deps * v02, "a1", "v2", "h_quad", "h_quad")
    el.add(eta_H * (1.0 - efrac) * drho / rho, "a1", "T", "h_quad", "h_quad")
    el.add(-eta_H * (dv02 - v02 * deps / eps), "a2", "v1", "h_cubic", "h_cubic")
    el.add(-eta_H * dv03, "a3", "v1", "h_cubic", "h_cubic")
    el.add(eta_H * (k2 * v02 + eps * k3 * v03), "a2", "v2", "h_cubic", "h_quad")
    el.add(eta_H * (k2 * v02 / eps + k3 * v03), "a3", "v3", "h_cubic", "h_quad")
    # viscous part
    el.add(-eta_H * IC * mu * (
        (drho / rho + 1.0 / eps) * deps / eps + (k2 / eps) ** 2 + k3**2
    ) / rho, "a1", "v1", "h_quad", "h_cubic")
    el.add(eta_H * IC * mu * (4.0 * drho / rho - deps / eps) / (3.0 * rho),
           "a1", "v1", "h_quad", "dh_cubic")
    el.add(eta_H * IC * mu * deps / (eps * rho), "a1", "v1", "dh_quad", "h_cubic")
    el.add(-4.0 * eta_H * IC * mu / (3.0 * rho), "a1", "v1", "dh_quad", "dh_cubic")
    el.add(eta_H * mu * 4.0 * (ddv01 + deps * (dv01 - v01 / eps) / eps) / (3.0 * rho**2),
           "a1", "rho", "h_quad", "h_quad")
    el.add(7.0 * eta_H * IC * mu * deps * k2 / (3.0 * eps * rho), "a1", "v2", "h_quad", "h_quad")
    el.add(eta_H * IC * mu * k3 * deps / (3.0 * eps * rho), "a1", "v3", "h_quad", "h_quad")
    el.add(-eta_H * IC * mu * k2 / (3.0 * rho), "a1", "v2", "h_quad", "dh_quad")
    el.add(-eta_H * IC * mu * k3 / (3.0 * rho), "a1", "v3", "h_quad", "dh_quad")
    el.add(2.0 * eta_H * IC * mu * k2 * deps / (eps**2 * rho), "a2", "v1", "h_cubic", "h_cubic")
    el.add(eta_H * IC * mu * k2 / (3.0 * eps * rho), "a2", "v1", "h_cubic", "dh_cubic")
    el.add(eta_H * IC * mu * k3 / (3.0 * rho), "a3", "v1", "h_cubic", "dh_cubic")
    el.add(-IC * eta_H * mu * (ddv02 + deps * (dv02 - v02 / eps) / eps) / rho**2,
           "a2", "rho", "h_cubic", "h_quad")
    el.add(-eta_H * IC * mu * (
        4.0 * k2**2 / (3.0 * eps) + eps * k3**2 + deps / eps
    ) / rho, "a2", "v2", "h_cubic", "h_quad")
    el.add(-eta_H * IC * mu * k2 * k3 / (3.0 * eps * rho), "a2", "v3", "h_cubic", "h_quad")
    el.add(-IC * eta_H * mu * (ddv03 + deps * dv03 / eps) / rho**2,
           "a3", "rho", "h_cubic", "h_quad")
    el.add(-eta_H * IC * mu * k2 * k3 / (3.0 * rho), "a3", "v2", "h_cubic", "h_quad")
    el.add(-eta_H * IC * mu * (
        (k2 / eps) ** 2 + 4.0 * k3**2 / 3.0 + drho * deps / (eps * rho)
    ) / rho, "a3", "v3", "h_cubic", "h_quad")
    el.add(eta_H * IC * mu * eps * drho / rho**2, "a2", "v2", "h_cubic", "dh_quad")
    el.add(eta_H * IC * mu * drho / rho**2, "a3", "v3", "h_cubic", "dh_quad")
    el.add(-eta_H * IC * mu * eps / rho, "a2", "v2", "dh_cubic", "dh_quad")
    el.add(-eta_H * IC * mu / rho, "a3", "v3", "dh_cubic", "dh_quad")
    el.add(eta_H * IC * mu * deps / (eps * rho), "a3", "v3", "dh_cubic", "h_quad")
    return el


# ------------------------------------------------------------ natural boundary terms
def add_natural_regular_terms(x, s, f):
    """smod_natural_bounds_regular.f08."""
    k2, k3 = s.k2, s.k3
    eps, _ = _eps(s, x)
    rho, T0 = f["rho0"], f["T0"]
    B01, B02, B03 = f["B01"], f["B02"], f["B03"]
    Gop_min = k3 * B02 - k2 * B03 / eps
    el = Elements(s.state_vector)
    el.add(T0, "v1", "rho", "h_cubic", "h_quad")
    el.add(rho, "v1", "T", "h_cubic", "h_quad")
    el.add(eps * Gop_min, "v1", "a1", "h_cubic", "h_quad")
    el.add(B03, "v1", "a2", "h_cubic", "dh_cubic")
    el.add(-eps * B02, "v1", "a3", "h_cubic", "dh_cubic")
    el.add(IC * eps * k3 * B01, "v2", "a1", "h_quad", "h_quad")
    el.add(-IC * k2 * B01, "v3", "a1", "h_quad", "h_quad")
    el.add(-IC * eps * B01, "v2", "a3", "h_quad", "dh_cubic")
    el.add(IC * B01, "v3", "a2", "h_quad", "dh_cubic")
    return el


def add_natural_flow_terms(x, s, f):
    """smod_natural_bounds_flow.f08."""
    el = Elements(s.state_vector)
    if not s.flow:
        return el
    rho, v01 = f["rho0"], f["v01"]
    el.add(-IC * rho * v01, "v1", "v1", "h_cubic", "h_cubic")
    el.add(-IC * rho * v01, "v3", "v3", "h_quad", "h_quad")
    el.add(-IC * rho * v01, "T", "T", "h_quad", "h_quad")
    return el


def add_natural_resistive_terms(x, s, f):
    """smod_natural_bounds_resistive.f08."""
    el = Elements(s.state_vector)
    if not s.resistivity:
        return el
    k2, k3 = s.k2, s.k3
    gamma_1 = s.gamma_1
    eps, deps = _eps(s, x)
    eta = f["eta"]
    B02, dB02, dB03 = f["B02"], f["dB02"], f["dB03"]
    drB02 = deps * B02 + eps * dB02
    el.add(2.0 * IC * gamma_1 * eta * (k3 * drB02 - k2 * dB03), "T", "a1", "h_quad", "h_quad")
    el.add(2.0 * IC * gamma_1 * eta * dB03, "T", "a2", "h_quad", "dh_cubic")
    el.add(-2.0 * IC * gamma_1 * eta * drB02, "T", "a3", "h_quad", "dh_cubic")
    el.add(-IC * eta * k2, "a2", "a1", "h_cubic", "h_quad")
    el.add(-IC * eta * eps * k3, "a3", "a1", "h_cubic", "h_quad")
    el.add(IC * eta, "a2", "a2", "h_cubic", "dh_cubic")
    el.add(IC * eta * eps, "a3", "a3", "h_cubic", "dh_cubic")
    return el


def add_natural_conduction_terms(x, s, f):
    """smod_natural_bounds_conduction.f08."""
    el = Elements(s.state_vector)
    if not s.conduction:
        return el
    k2, k3 = s.k2, s.k3
    gamma_1 = s.gamma_1
    eps, deps = _eps(s, x)
    dT0 = f["dT0"]
    dkappa_para_dT = f["dtcparadT"]
    kappa_perp = f["tcperp"]
    dkappa_perp_drho = f["dtcperpdrho"]
    dkappa_perp_dT = f["dtcperpdT"]
    el.add(IC * gamma_1 * dT0 * dkappa_perp_drho, "T", "rho", "h_quad", "h_quad")
    el.add(gamma_1 * (-deps * IC * kappa_perp / eps + IC * dT0 * dkappa_perp_dT),
           "T", "T", "h_quad", "h_quad")
    el.add(IC * gamma_1 * kappa_perp, "T", "T", "h_quad", "dh_quad")
    if not s.has_bfield:
        return el
    B0 = _B0(f)
    B01, B02, B03 = f["B01"], f["B02"], f["B03"]
    dkappa_perp_dB2 = f["dtcperpdB2"]
    Gop_min = k3 * B02 - k2 * B03 / eps
    Fop = k2 * B02 / eps + k3 * B03
    Kp = f["tcprefactor"]
    Kp_plus = Kp + dkappa_perp_dB2
    Kp_plusplus = dkappa_perp_dB2 - (B01**2 * Kp_plus / B0**2)
    el.add(-IC * gamma_1 * dT0 * dkappa_perp_drho * B01**2 / B0**2,
           "T", "rho", "h_quad", "h_quad")
    el.add(gamma_1 * (
        -B01 * Kp * (2.0 * (deps / eps) * IC * B01 + 3.0 * Fop)
        + IC * dT0 * (B01**2 * dkappa_para_dT / B0**2 - dkappa_perp_dT * B01**2 / B0**2)
    ), "T", "T", "h_quad", "h_quad")
    el.add(2.0 * IC * gamma_1 * eps * dT0 * Gop_min * Kp_plusplus, "T", "a1", "h_quad", "h_quad")
    el.add(IC * gamma_1 * 2.0 * B01**2 * Kp, "T", "T", "h_quad", "dh_quad")
    el.add(2.0 * gamma_1 * k3 * dT0 * B01 * Kp_plusplus, "T", "a2", "h_quad", "h_cubic")
    el.add(-2.0 * gamma_1 * k2 * dT0 * B01 * Kp_plusplus, "T", "a3", "h_quad", "h_cubic")
    el.add(2.0 * IC * gamma_1 * dT0 * B03 * Kp_plusplus, "T", "a2", "h_quad", "dh_cubic")
    el.add(-2.0 * IC * gamma_1 * dT0 * eps * B02 * Kp_plusplus, "T", "a3", "h_quad", "dh_cubic")
    return el


def add_natural_viscosity_terms(x, s, f):
    """smod_natural_bounds_viscosity.f08."""
    el = Elements(s.state_vector)
    if not s.viscosity:
        return el
    k2, k3 = s.k2, s.k3
    gamma_1 = s.gamma_1
    eps, deps = _eps(s, x)
    mu = s.viscosity_value
    dv01, dv03 = f["dv01"], f["dv03"]
    one = np.ones_like(x)
    el.add(-IC * mu * deps / eps, "v1", "v1", "h_cubic", "h_cubic")
    el.add(4.0 * IC * mu / 3.0 * one, "v1", "v1", "h_cubic", "dh_cubic")
    el.add(-IC * mu * k2 / 3.0 * one, "v1", "v2", "h_cubic", "h_quad")
    el.add(-IC * mu * k3 / 3.0 * one, "v1", "v3", "h_cubic", "h_quad")
    el.add(IC * mu * eps, "v2", "v2", "h_quad", "dh_quad")
    el.add(IC * mu * one, "v3", "v3", "h_quad", "dh_quad")
    el.add(-IC * mu * deps / eps, "v3", "v3", "h_quad", "h_quad")
    if s.viscous_heating and not s.incompressible:
        el.add(2.0 * IC * gamma_1 * mu * dv03, "T", "v3", "h_quad", "h_quad")
        el.add(2.0 * gamma_1 * mu * dv01, "T", "v1", "h_quad", "h_cubic")
    return el


def add_natural_hall_terms(x, s, f):
    """smod_natural_bounds_hall.f08 (A-matrix part)."""
    el = Elements(s.state_vector)
    if not s.hall or not s.viscosity:
        return el
    eps, deps = _eps(s, x)
    rho = f["rho0"]
    eta_H = f["hallfactor"]
    mu = s.viscosity_value
    el.add(-eta_H * IC * mu * deps / (eps * rho), "a1", "v1", "h_quad", "h_cubic")
    el.add(4.0 * eta_H * IC * mu / (3.0 * rho), "a1", "v1", "h_quad", "dh_cubic")
    el.add(eta_H * IC * mu * eps / rho, "a2", "v2", "h_cubic", "dh_quad")
    el.add(eta_H * IC * mu / rho, "a3", "v3", "h_cubic", "dh_quad")
    el.add(-eta_H * IC * mu * deps / (eps * rho), "a3", "v3", "h_cubic", "h_quad")
    return el


def add_natural_hall_Bterms(x, s, f):
    """smod_natural_bounds_hall.f08 (B-matrix part)."""
    el = Elements(s.state_vector)
    if not s.hall or not s.electron_inertia:
        return el
    k2, k3 = s.k2, s.k3
    eps, _ = _eps(s, x)
    rho = f["rho0"]
    eta_e = f["inertiafactor"]
    el.add(eta_e * k2 / rho, "a2", "a1", "h_cubic", "h_quad")
    el.add(eta_e * k3 * eps / rho, "a3", "a1", "h_cubic", "h_quad")
    el.add(-eta_e / rho, "a2", "a2", "h_cubic", "dh_cubic")
    el.add(-eta_e * eps / rho, "a3", "a3", "h_cubic", "dh_cubic")
    return el


# ------------------------------------------------------------------- matrix storage
class BlockTriMatrix:
    """Assembled matrix in block-tridiagonal form plus the reference's node structure.

    ``blocks[b, 0|1|2]`` = sub / diagonal / super ``d x d`` block of block row ``b``
    (d = dim_subblock).  ``mask`` marks entries that own a node in the reference's
    linked lists, ``phase`` the insertion phase that created the node (0: element
    b-1, 1: element b, 2: natural boundary, 3: essential boundary) which, with the
    column index, gives the reference's per-row insertion order.
    """

    def __init__(self, gridpts: int, dsub: int, label: str):
        self.G, self.d, self.label = gridpts, dsub, label
        self.blocks = np.zeros((gridpts, 3, dsub, dsub), dtype=np.complex128)
        self.mask = np.zeros((gridpts, 3, dsub, dsub), dtype=bool)
        self.phase = np.zeros((gridpts, 3, dsub, dsub), dtype=np.int8)

    @property
    def n(self):
        return self.G * self.d

    def add(self, b, which, contrib, phase):
        """matrix_t%add_element semantics for whole blocks (drop rule per contribution)."""
        keep = ~((np.abs(contrib.real) <= DP_LIMIT) & (np.abs(contrib.imag) <= DP_LIMIT))
        tgt = self.blocks[b, which]
        msk = self.mask[b, which]
        new = keep & ~msk
        tgt[keep] += contrib[keep]
        self.phase[b, which][new] = phase
        msk |= keep

    def add_quadblock(self, e, quad, phase_lo, phase_hi):
        """Insert a full quadblock whose top-left corner is block row ``e``."""
        d = self.d
        self.add(e, 1, quad[:d, :d], phase_lo)
        self.add(e, 2, quad[:d, d:], phase_lo)
        self.add(e + 1, 0, quad[d:, :d], phase_hi)
        self.add(e + 1, 1, quad[d:, d:], phase_hi)

    def _locate(self, row, col):
        b, i = divmod(row, self.d)
        bc, j = divmod(col, self.d)
        return b, bc - b + 1, i, j

    def delete(self, row, col):
        b, w, i, j = self._locate(row, col)
        if 0 <= w <= 2:
            self.blocks[b, w, i, j] = 0.0
            self.mask[b, w, i, j] = False

    def to_dense(self):
        n, d = self.n, self.d
        out = np.zeros((n, n), dtype=np.complex128)
        for b in range(self.G):
            for w in range(3):
                bc = b + w - 1
                if 0 <= bc < self.G:
                    out[b * d:(b + 1) * d, bc * d:(bc + 1) * d] = self.blocks[b, w]
        return out

    def to_band(self, kl=None, ku=None):
        """LAPACK general band: a_ij at AB[ku + i - j, j] (mod_transform_matrix.f08:80-106)."""
        d, n = self.d, self.n
        kl = 2 * d - 1 if kl is None else kl
        ku = 2 * d - 1 if ku is None else ku
        ab = np.zeros((kl + ku + 1, n), dtype=np.complex128)
        ii = np.arange(d)[:, None]
        jj = np.arange(d)[None, :]
        for w in range(3):
            lo, hi = max(0, 1 - w), min(self.G, self.G + 1 - w)
            if hi <= lo:
                continue
            b = np.arange(lo, hi)
            rows = (b * d)[:, None, None] + ii[None]
            cols = ((b + w - 1) * d)[:, None, None] + jj[None]
            r = np.broadcast_to(rows, (len(b), d, d))
            c = np.broadcast_to(cols, (len(b), d, d))
            ab[ku + r - c, c] = self.blocks[lo:hi, w]
        return ab

    def to_coo(self):
        """(rows, cols, vals) 1-based in the reference's output order
        (src/dataIO/mod_output.f08:489-506: rows ascending, list order per row)."""
        rows, cols, vals = [], [], []
        d = self.d
        for b in range(self.G):
            for i in range(d):
                ent = []
                for w in range(3):
                    for j in np.nonzero(self.mask[b, w, i])[0]:
                        col = (b + w - 1) * d + j
                        ent.append((int(self.phase[b, w, i, j]), col, self.blocks[b, w, i, j]))
                ent.sort(key=lambda t: (t[0], t[1]))
                for _, col, val in ent:
                    rows.append(b * d + i + 1)
                    cols.append(col + 1)
                    vals.append(val)
        return np.array(rows), np.array(cols), np.array(vals)

    def matvec(self, x):
        d = self.d
        xb = x.reshape(self.G, d)
        y = np.einsum("bij,bj->bi", self.blocks[:, 1], xb)
        y[1:] += np.einsum("bij,bj->bi", self.blocks[1:, 0], xb[:-1])
        y[:-1] += np.einsum("bij,bj->bi", self.blocks[:-1, 2], xb[1:])
        return y.reshape(-1)


# ------------------------------------------------------------------ build_matrices
def complete_fields(fields: dict, npts: int) -> dict:
    out = {}
    for name in FIELD_NAMES:
        val = fields.get(name)
        if val is None:
            out[name] = np.zeros(npts)
        else:
            out[name] = np.broadcast_to(np.asarray(val, dtype=np.float64), (npts,)).copy()
    unknown = set(fields) - set(FIELD_NAMES)
    if unknown:
        raise KeyError(f"unknown field(s): {sorted(unknown)}")
    return out


def element_quadblocks(s: Settings, grid, gauss_grid, fields):
    """Quadblocks of all elements: returns (QA, QB) with shape (dimq, dimq, G-1)."""
    n = s.gridpts - 1
    dimq, dsub = s.dim_quadblock, s.dim_subblock
    QA = np.zeros((dimq, dimq, n), dtype=np.complex128)
    QB = np.zeros((dimq, dimq, n), dtype=np.complex128)
    x_left, x_right = grid[:-1], grid[1:]
    for j in range(4):
        x = gauss_grid[j::4]
        f = {k: v[j::4] for k, v in fields.items()}
        w = s.gauss_weights[j]
        spl = _splines(x, x_left, x_right)
        add_to_quadblock(QB, add_bmatrix_terms(x, s, f), w, dsub, spl)
        add_to_quadblock(QA, add_regular_matrix_terms(x, s, f), w, dsub, spl)
        if s.flow:
            add_to_quadblock(QA, add_flow_matrix_terms(x, s, f), w, dsub, spl)
        if s.resistivity:
            add_to_quadblock(QA, add_resistive_matrix_terms(x, s, f), w, dsub, spl)
        if s.cooling or s.heating:
            add_to_quadblock(QA, add_heatloss_matrix_terms(x, s, f), w, dsub, spl)
        if s.conduction:
            add_to_quadblock(QA, add_conduction_matrix_terms(x, s, f), w, dsub, spl)
        if s.viscosity:
            add_to_quadblock(QA, add_viscosity_matrix_terms(x, s, f), w, dsub, spl)
        if s.hall:
            add_to_quadblock(QA, add_hall_matrix_terms(x, s, f), w, dsub, spl)
            add_to_quadblock(QB, add_hall_bmatrix_terms(x, s, f), w, dsub, spl)
    dx = x_right - x_left
    QB *= dx
    QA *= dx
    return QA, QB


def _scatter(mat: BlockTriMatrix, Q):
    """mod_matrix_manager.f08:251-259, vectorised over elements."""
    d = mat.d
    n = Q.shape[2]
    Qt = np.moveaxis(Q, 2, 0)  # (n, dimq, dimq)
    keep = ~((np.abs(Qt.real) <= DP_LIMIT) & (np.abs(Qt.imag) <= DP_LIMIT))
    Qk = np.where(keep, Qt, 0.0)
    e = np.arange(n)
    # phase 0 contributions (element b-1 seen from block row b): bottom half
    mat.blocks[e + 1, 0] += Qk[:, d:, :d]
    mat.mask[e + 1, 0] |= keep[:, d:, :d]
    mat.blocks[e + 1, 1] += Qk[:, d:, d:]
    mat.mask[e + 1, 1] |= keep[:, d:, d:]
    # phase 1 contributions (element b): top half; phase recorded where node is new
    new = keep[:, :d, :d] & ~mat.mask[e, 1]
    mat.phase[e, 1] = np.where(new, 1, mat.phase[e, 1])
    mat.blocks[e, 1] += Qk[:, :d, :d]
    mat.mask[e, 1] |= keep[:, :d, :d]
    mat.phase[e, 2] = 1
    mat.blocks[e, 2] += Qk[:, :d, d:]
    mat.mask[e, 2] |= keep[:, :d, d:]


def _natural_quadblock(s: Settings, grid, gauss_grid, fields, edge, label):
    """smod_natural_boundaries.f08:103-222 — one edge quadblock (dimq, dimq)."""
    dimq, dsub = s.dim_quadblock, s.dim_subblock
    quad = np.zeros((dimq, dimq, 1), dtype=np.complex128)
    if edge == "left":
        lo, hi = grid[0:1], grid[1:2]
        pos, idx, weight = lo, 0, -1.0
    else:
        lo, hi = grid[-2:-1], grid[-1:]
        pos, idx, weight = hi, len(gauss_grid) - 1, 1.0
    spl = _splines(pos, lo, hi)
    # equilibrium at the first/last *Gaussian* point (:113, :155)
    x = gauss_grid[idx:idx + 1]
    f = {k: v[idx:idx + 1] for k, v in fields.items()}
    if label == "A":
        for proc in (add_natural_regular_terms, add_natural_flow_terms,
                     add_natural_resistive_terms, add_natural_conduction_terms,
                     add_natural_viscosity_terms, add_natural_hall_terms):
            add_to_quadblock(quad, proc(x, s, f), weight, dsub, spl)
    else:
        add_to_quadblock(quad, add_natural_hall_Bterms(x, s, f), weight, dsub, spl)
    return quad[:, :, 0]


def _subblock_index(variables, s: Settings, odd: bool, edge: str):
    """mod_get_indices.f08:96-131 (1-based indices inside the edge quadblock)."""
    idxs = [2 * (s.state_vector.index(v) + 1) for v in variables if v in s.state_vector]
    if odd:
        idxs = [i - 1 for i in idxs]
    if edge == "right":
        idxs = [i + s.dim_subblock for i in idxs]
    return idxs


def essential_indices(s: Settings, edge: str):
    """Quadblock-local 1-based indices zeroed at an edge, in the reference's order
    (smod_essential_boundaries.f08:12-157; flags mod_boundary_manager.f08:87-107)."""
    def is_zero(v):
        return abs(v) <= DP_LIMIT

    cubic = ["v1"]
    if s.boundary_type == "wall":
        cubic += ["a2", "a3"]
    elif s.boundary_type == "wall_weak":
        if not is_zero(s.k2):
            cubic += ["a3"]
        if not is_zero(s.k3):
            cubic += ["a2"]
    else:
        raise ValueError(f"unknown boundary_type {s.boundary_type}")
    apply_T = s.perpendicular_conduction
    noslip_right = s.viscosity
    noslip_left = s.viscosity and (s.coaxial or s.geometry == "Cartesian")
    out = []
    if edge == "left":
        out += _subblock_index(["rho", "v2", "v3", "T", "a1"], s, True, "left")
        out += _subblock_index(cubic, s, True, "left")
        if apply_T:
            out += _subblock_index(["T"], s, False, "left")
        if noslip_left:
            out += _subblock_index(["v2", "v3"], s, False, "left")
    else:
        out += _subblock_index(cubic, s, True, "right")
        if apply_T:
            out += _subblock_index(["T"], s, False, "right")
        if noslip_right:
            out += _subblock_index(["v2", "v3"], s, False, "right")
    return out


def _apply_essential(mat: BlockTriMatrix, s: Settings, edge: str):
    dimq = s.dim_quadblock
    shift = 0 if edge == "left" else mat.n - dimq
    for idx in essential_indices(s, edge):
        g = shift + idx - 1  # 0-based global
        for k in range(shift, shift + dimq):
            mat.delete(g, k)
            mat.delete(k, g)
        if mat.label == "B":  # A gets 0 which add_element drops (:189-205)
            b, w, i, j = mat._locate(g, g)
            mat.blocks[b, w, i, j] += 1.0
            if not mat.mask[b, w, i, j]:
                mat.phase[b, w, i, j] = 3
            mat.mask[b, w, i, j] = True


def build_matrices(s: Settings, grid, gauss_grid, fields):
    """Restatement of ``build_matrices`` + ``apply_boundary_conditions``.

    Returns ``(A, B)`` as :class:`BlockTriMatrix`.
    """
    grid = np.asarray(grid, dtype=np.float64)
    gauss_grid = np.asarray(gauss_grid, dtype=np.float64)
    assert len(grid) == s.gridpts and len(gauss_grid) == 4 * (s.gridpts - 1)
    fields = complete_fields(fields, len(gauss_grid))
    QA, QB = element_quadblocks(s, grid, gauss_grid, fields)
    A = BlockTriMatrix(s.gridpts, s.dim_subblock, "A")
    B = BlockTriMatrix(s.gridpts, s.dim_subblock, "B")
    _scatter(B, QB)
    _scatter(A, QA)
    last = s.gridpts - 2
    # order: mod_boundary_manager.f08:72-83
    for mat, edge in ((B, "left"), (A, "left"), (B, "right"), (A, "right")):
        quad = _natural_quadblock(s, grid, gauss_grid, fields, edge, mat.label)
        mat.add_quadblock(0 if edge == "left" else last, quad, 2, 2)
        _apply_essential(mat, s, edge)
    return A, B
