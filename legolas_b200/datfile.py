"""Row N2 of the scope table: the Legolas 2.x datfile, written from device-resident results.

Host-side mirror of ``create_datfile`` (src/dataIO/mod_output.f08:56-105): same byte layout
(stream access, native endianness, no record markers), so pylbo's reader
(post_processing/pylbo/utilities/datfiles/file_reader.py:52-157, header.py:53-338) and the
reference's regression tooling read a GPU run like a CPU run.  What the reference takes from its
linked-list ``matrix_t`` comes from the device here:

* ``write_matrices``  -> ``lgpu_export_coo`` (triplets already in the reference's insertion order,
  B before A, B real; mod_output.f08:476-508),
* ``write_residuals`` -> ``lgpu_residuals`` (mod_output.f08:445-473, 511-545),
* ``write_eigenfunctions`` -> ``lgpu_eigenfunctions`` (mod_output.f08:403-415).

Sections are written in the reference's order: version tag, header (physics type, grid, io,
solver, equilibrium, units, physics, parameters, background names), eigenvalues, grids,
background arrays, eigenfunctions, eigenvectors, residuals, matrices.  Derived eigenfunctions
(``write_derived_eigenfunctions``) are not produced by this library: the flag is written as
false, as a reference run with ``write_derived_eigenfunctions = .false.`` does.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, Mapping, Optional, Sequence

import numpy as np

from .api import STATE_VECTORS, Context, LegolasError, Settings

LEGOLAS_VERSION = "2.0.6"          # src/mod_version.f08:17 (character(len=10))
STR_LEN, STR_LEN_ARR = 500, 16     # src/mod_global_variables.f08:14-16
GAUSSIAN_NODES = (-0.861136311594053, -0.339981043584856, 0.339981043584856, 0.861136311594053)
GAUSSIAN_WEIGHTS = (0.347854845137454, 0.652145154862546, 0.652145154862546, 0.347854845137454)

# write_units_info (mod_output.f08:213-243): names in file order
UNIT_NAMES = ("unit_length", "unit_time", "unit_density", "unit_velocity", "unit_temperature",
              "unit_pressure", "unit_magneticfield", "unit_numberdensity", "unit_mass",
              "mean_molecular_weight", "unit_resistivity", "unit_lambdaT", "unit_conduction")
# write_parameters (mod_output.f08:288-318)
PARAMETER_NAMES = ("k2", "k3", "cte_rho0", "cte_T0", "cte_p0", "cte_B01", "cte_B02", "cte_B03", "Bth0",
                   "Bz0", "cte_v02", "cte_v03", "p1", "p2", "p3", "p4", "p5", "p6", "p7", "p8", "alpha",
                   "beta", "delta", "theta", "tau", "lambda", "nu", "r0", "rc", "rj", "V", "j0", "g",
                   "electronfraction", "viscosity_value")
# write_background_names / write_background_data (mod_output.f08:321-400): file order, and the
# slot of the C ABI (api.FIELD_NAMES) or derived quantity each one comes from
BACKGROUND_NAMES = ("rho0", "drho0", "T0", "dT0", "ddT0", "B01", "B02", "B03", "dB02", "dB03", "ddB02",
                    "ddB03", "B0", "v01", "v02", "v03", "dv01", "dv02", "dv03", "ddv01", "ddv02", "ddv03",
                    "L0", "dLdT", "dLdrho", "lambdaT", "dlambdadT", "H0", "dHdT", "dHdrho", "kappa_para",
                    "kappa_perp", "dkappa_para_dT", "dkappa_para_dr", "dkappa_perp_drho",
                    "dkappa_perp_dT", "dkappa_perp_dB2", "dkappa_perp_dr", "eta", "detadT", "detadr",
                    "gravity", "Hall", "inertia")
_FROM_FIELD = {"kappa_para": "tcpara", "kappa_perp": "tcperp", "dkappa_para_dT": "dtcparadT",
               "dkappa_perp_drho": "dtcperpdrho", "dkappa_perp_dT": "dtcperpdT",
               "dkappa_perp_dB2": "dtcperpdB2", "gravity": "g0", "Hall": "hallfactor",
               "inertia": "inertiafactor"}


@dataclass
class IoSettings:
    """io_settings_t (src/settings/mod_io_settings.f08): the output switches."""

    write_matrices: bool = False
    write_eigenvectors: bool = False
    write_residuals: bool = False
    write_eigenfunctions: bool = False
    write_background: bool = True
    write_ef_subset: bool = False
    ef_subset_radius: float = float("nan")
    ef_subset_center: complex = complex(float("nan"), float("nan"))


@dataclass
class RunInfo:
    """Header entries the hot path never reads but the datfile carries (host settings)."""

    grid_start: Optional[float] = None            # default: first / last base grid point
    grid_end: Optional[float] = None
    equilibrium_type: str = "user_defined"
    cgs: bool = True
    units: Mapping[str, float] = field(default_factory=dict)        # UNIT_NAMES -> value (missing: NaN)
    parameters: Mapping[str, float] = field(default_factory=dict)   # PARAMETER_NAMES -> value (missing: NaN)
    cooling_curve: str = "jc_corona"              # physics defaults of the reference
    interpolation_points: int = 4000
    fixed_resistivity: bool = False
    fixed_tc_para: bool = False
    fixed_tc_perp: bool = False
    hall_uses_substitution: bool = True
    extra_background: Mapping[str, np.ndarray] = field(default_factory=dict)   # e.g. lambdaT, H0, dkappa_*_dr


class _Writer:
    def __init__(self, fh):
        self.fh = fh

    def raw(self, fmt, *vals):
        self.fh.write(struct.pack("=" + fmt, *vals))

    def i(self, *vals):
        self.raw(f"{len(vals)}i", *[int(v) for v in vals])

    def b(self, val):          # Fortran logical: 4 bytes
        self.raw("i", 1 if val else 0)

    def d(self, *vals):
        self.raw(f"{len(vals)}d", *[float(v) for v in vals])

    def z(self, val):
        val = complex(val)
        self.raw("2d", val.real, val.imag)

    def s(self, text, length=None):
        data = text.encode("ascii")
        if length is not None:
            data = data[:length].ljust(length, b" ")
        self.fh.write(data)

    def ls(self, text):        # write(dat_fh) len(x), x
        self.i(len(text))
        self.s(text)

    def arr(self, a, dtype):
        self.fh.write(np.ascontiguousarray(a, dtype=dtype).tobytes())


def ef_grid(base_grid: np.ndarray) -> np.ndarray:
    """grid%ef_grid (src/mod_grid.f08:142-157): every base grid point and every interval midpoint."""
    base_grid = np.asarray(base_grid, dtype=np.float64)
    out = np.empty(2 * base_grid.size - 1)
    out[0::2] = base_grid
    out[1::2] = 0.5 * (base_grid[:-1] + base_grid[1:])
    return out


def select_ef_subset(omega: np.ndarray, io: IoSettings):
    """ef_written_flags / ef_written_idxs (src/eigenfunctions/mod_eigenfunctions.f08:150-176):
    all eigenvalues, or those within ef_subset_radius of ef_subset_center; indices are 1-based."""
    omega = np.asarray(omega, dtype=np.complex128)
    if io.write_ef_subset:
        flags = np.abs(omega - io.ef_subset_center) <= io.ef_subset_radius
    else:
        flags = np.ones(omega.size, dtype=bool)
    return flags, (np.nonzero(flags)[0] + 1).astype(np.int32)


def background_arrays(settings: Settings, gauss_grid: np.ndarray, fields: Mapping[str, np.ndarray],
                      extra: Mapping[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """The 44 arrays of write_background_data from the sampled fields of the C ABI.  B0 is
    sqrt(B01^2 + B02^2 + B03^2) (mod_bg_magnetic.f08 get_B0); quantities the hot path has no
    slot for (lambdaT, dlambdadT, H0, dHdT, dHdrho, dkappa_para_dr, dkappa_perp_dr) come from
    `extra` or are written as zeros, which is what the reference writes when that physics is off."""
    n = np.asarray(gauss_grid).size

    def get(name):
        val = fields.get(name)
        if val is None:
            return np.zeros(n)
        return np.ascontiguousarray(np.broadcast_to(np.asarray(val, dtype=np.float64), (n,)))

    out = {}
    for name in BACKGROUND_NAMES:
        if name in extra:
            out[name] = np.ascontiguousarray(np.broadcast_to(np.asarray(extra[name], dtype=np.float64), (n,)))
        elif name == "B0":
            out[name] = np.sqrt(get("B01") ** 2 + get("B02") ** 2 + get("B03") ** 2)
        else:
            out[name] = get(_FROM_FIELD.get(name, name))
    return out


def create_datfile(path, settings: Settings, base_grid: np.ndarray, gauss_grid: np.ndarray,
                   fields: Mapping[str, np.ndarray], eigenvalues: Sequence[complex],
                   ctx: Optional[Context] = None, eigenvectors: Optional[np.ndarray] = None,
                   io: Optional[IoSettings] = None, info: Optional[RunInfo] = None) -> str:
    """Write `path` in the Legolas 2.0.6 datfile layout.

    `ctx` holds the assembled matrices of this run (needed for write_matrices, write_residuals and
    write_eigenfunctions); `eigenvectors` is (dim_matrix, nev) in the reference's numbering (needed
    for write_eigenvectors, write_residuals, write_eigenfunctions)."""
    io = io or IoSettings()
    info = info or RunInfo()
    omega = np.ascontiguousarray(eigenvalues, dtype=np.complex128)
    base_grid = np.ascontiguousarray(base_grid, dtype=np.float64)
    gauss_grid = np.ascontiguousarray(gauss_grid, dtype=np.float64)
    sv = settings.solvers
    need_vectors = io.write_eigenvectors or io.write_residuals or io.write_eigenfunctions
    if need_vectors:
        if eigenvectors is None:
            raise LegolasError("create_datfile: eigenvectors requested but not supplied")
        eigenvectors = np.asarray(eigenvectors, dtype=np.complex128)
        if eigenvectors.shape != (settings.dim_matrix, omega.size):
            raise LegolasError(f"create_datfile: eigenvectors have shape {eigenvectors.shape}, expected "
                               f"{(settings.dim_matrix, omega.size)}")
    if (io.write_matrices or io.write_residuals or io.write_eigenfunctions) and ctx is None:
        raise LegolasError("create_datfile: a Context with the assembled matrices is required")
    state_vector = STATE_VECTORS[settings.physics_type]
    nodes = GAUSSIAN_NODES if settings.gauss_nodes is None else tuple(settings.gauss_nodes)
    weights = GAUSSIAN_WEIGHTS if settings.gauss_weights is None else tuple(settings.gauss_weights)

    with open(path, "wb") as fh:
        w = _Writer(fh)
        w.s("legolas_version")
        w.s(LEGOLAS_VERSION, 10)
        w.i(STR_LEN, STR_LEN_ARR)
        # ---- write_physics_type_info
        w.i(settings.nb_eqs)
        w.ls(settings.physics_type)
        w.i(STR_LEN_ARR, len(state_vector))     # character(len=str_len_arr) (mod_settings.f08:18)
        for name in state_vector:
            w.s(name, STR_LEN_ARR)
        nb = settings.nb_eqs     # dims (src/settings/mod_dims.f08:36-44)
        w.i(2, 2 * nb, 4 * nb, settings.dim_matrix)   # integralblock is 2 for every physics type
        # ---- write_grid_info
        w.ls(settings.geometry)
        w.i(settings.gridpts, 4 * (settings.gridpts - 1), 2 * settings.gridpts - 1)
        w.i(4)
        w.d(*nodes)
        w.d(*weights)
        w.d(base_grid[0] if info.grid_start is None else info.grid_start)
        w.d(base_grid[-1] if info.grid_end is None else info.grid_end)
        # ---- write_io_info
        for flag in (io.write_matrices, io.write_eigenvectors, io.write_residuals, io.write_eigenfunctions,
                     False, io.write_ef_subset):
            w.b(flag)
        w.d(io.ef_subset_radius)
        w.z(io.ef_subset_center)
        # ---- write_solver_info
        w.ls(sv.solver)
        w.ls(sv.arpack_mode)
        w.i(sv.number_of_eigenvalues)
        w.ls(sv.which_eigenvalues)
        w.i(sv.ncv, sv.maxiter)
        w.z(sv.sigma)
        w.d(sv.tolerance)
        # ---- write_equilibrium_info
        w.ls(info.equilibrium_type)
        w.ls(settings.boundary_type)
        # ---- write_units_info
        w.i(len(UNIT_NAMES))
        w.b(info.cgs)
        for name in UNIT_NAMES:
            w.ls(name)
            w.d(info.units.get(name, float("nan")))
        # ---- write_physics_info
        # set_incompressible overwrites gamma (src/settings/mod_physics_settings.f08:90-94)
        w.d(1.0e12 if settings.incompressible else settings.gamma)
        w.b(settings.incompressible)
        w.b(settings.flow)
        w.b(settings.cooling)
        w.ls(info.cooling_curve)
        w.i(info.interpolation_points)
        w.b(settings.gravity)
        w.b(settings.resistivity)
        w.b(info.fixed_resistivity)
        w.b(settings.viscosity)
        w.b(settings.viscous_heating)
        w.b(settings.conduction)
        # conduction%is_enabled() = parallel .or. perpendicular (src/settings/mod_physics_settings.f08)
        w.b(settings.conduction if settings.parallel_conduction is None else settings.parallel_conduction)
        w.b(info.fixed_tc_para)
        w.b(settings.perpendicular_conduction)
        w.b(info.fixed_tc_perp)
        w.b(settings.hall)
        w.b(info.hall_uses_substitution)
        w.b(settings.electron_inertia)
        # ---- write_parameters
        params = dict(info.parameters)
        params.setdefault("k2", settings.k2)
        params.setdefault("k3", settings.k3)
        params.setdefault("electronfraction", settings.electron_fraction)
        params.setdefault("viscosity_value", settings.viscosity_value)
        unknown = set(params) - set(PARAMETER_NAMES)
        if unknown:
            raise LegolasError(f"create_datfile: unknown parameter(s) {sorted(unknown)}")
        w.i(len(PARAMETER_NAMES), STR_LEN_ARR)
        for name in PARAMETER_NAMES:
            w.s(name, STR_LEN_ARR)
            w.d(params.get(name, float("nan")))
        # ---- write_background_names
        if io.write_background:
            w.i(len(BACKGROUND_NAMES), STR_LEN_ARR)
            for name in BACKGROUND_NAMES:
                w.s(name, STR_LEN_ARR)
        else:
            w.i(0, 0)
        # ---- data blocks (create_datfile, mod_output.f08:84-102)
        w.i(omega.size)
        w.arr(omega, np.complex128)
        w.arr(base_grid, np.float64)
        w.arr(gauss_grid, np.float64)
        if io.write_background:
            for name, arr in background_arrays(settings, gauss_grid, fields, info.extra_background).items():
                w.arr(arr, np.float64)
        if io.write_eigenfunctions:
            grid_ef = ef_grid(base_grid)
            flags, idxs = select_ef_subset(omega, io)
            w.i(grid_ef.size)
            w.arr(grid_ef, np.float64)
            w.i(flags.size)
            w.arr(flags.astype(np.int32), np.int32)
            w.i(idxs.size)
            w.arr(idxs, np.int32)
            efs = ctx.eigenfunctions(eigenvectors, idxs) if idxs.size else {}
            for name in state_vector:     # (ef_gridpts, nb_written) column-major per variable
                block = efs[name] if idxs.size else np.zeros((grid_ef.size, 0), dtype=np.complex128)
                w.arr(np.asfortranarray(block).T, np.complex128)
        if io.write_eigenvectors:
            w.i(eigenvectors.shape[0], eigenvectors.shape[1])
            w.arr(np.asfortranarray(eigenvectors).T, np.complex128)
        if io.write_residuals:
            res = ctx.residuals(omega, eigenvectors)
            w.i(res.size)
            w.arr(res, np.float64)
        if io.write_matrices:
            rb, cb, vb = ctx.export_coo("B")
            ra, ca, va = ctx.export_coo("A")
            w.i(rb.size)
            w.i(ra.size)
            rec_b = np.empty(rb.size, dtype=np.dtype([("r", "<i4"), ("c", "<i4"), ("v", "<f8")]))
            rec_b["r"], rec_b["c"], rec_b["v"] = rb, cb, np.real(vb)   # B is real in the file
            fh.write(rec_b.tobytes())
            rec_a = np.empty(ra.size, dtype=np.dtype([("r", "<i4"), ("c", "<i4"), ("re", "<f8"), ("im", "<f8")]))
            rec_a["r"], rec_a["c"], rec_a["re"], rec_a["im"] = ra, ca, np.real(va), np.imag(va)
            fh.write(rec_a.tobytes())
    return str(path)
