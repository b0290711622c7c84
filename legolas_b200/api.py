"""Host-side mirror of the reference interface for the hot path.

The reference host is Fortran (no compiler in this image), so the layer above the C ABI is
mirrored here with the reference's names and argument meaning:

  build_matrices(settings, grid, background, physics) ... src/matrices/mod_matrix_manager.f08:138-153
  solve_evp(matrix_A, matrix_B, settings, omega, vr) ..... src/solvers/mod_solvers.f08:92-123
  new_arpack_config(evpdim, mode, bmat, solver_settings) . src/solvers/arnoldi/mod_arpack_type.f08:74-102
  solver defaults ........................................ src/settings/mod_solver_settings.f08:30-44

All arithmetic happens in liblegolas_b200.so on the GPU (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

from . import _lib
from ._lib import CArnoldi, CSettings, CStats, LgpuError, N_FIELDS

FIELD_NAMES = (
    "rho0", "drho0", "T0", "dT0", "ddT0",
    "B01", "B02", "dB02", "ddB02", "B03", "dB03", "ddB03",
    "v01", "dv01", "ddv01", "v02", "dv02", "ddv02", "v03", "dv03", "ddv03",
    "g0", "eta", "detadT", "detadr", "L0", "dLdT", "dLdrho",
    "tcpara", "dtcparadT", "tcperp", "dtcperpdrho", "dtcperpdT", "dtcperpdB2",
    "tcprefactor", "dtcprefactordr", "hallfactor", "inertiafactor",
)
assert len(FIELD_NAMES) == N_FIELDS

ZLARNV_SEED = (2022, 9, 30, 179)   # mod_arpack_type.f08:206
PHYSICS_TYPES = {"mhd": 0, "hd": 1, "hd-1d": 2}
# src/settings/mod_settings.f08:69-86
STATE_VECTORS = {"mhd": ("rho", "v1", "v2", "v3", "T", "a1", "a2", "a3"),
                 "hd": ("rho", "v1", "v2", "v3", "T"), "hd-1d": ("rho", "v1", "T")}
GEOMETRIES = {"Cartesian": 0, "cylindrical": 1}
BOUNDARY_TYPES = {"wall": 0, "wall_weak": 1}
ALLOWED_WHICH = ("LM", "SM", "LR", "SR", "LI", "SI")
MAX_NCV = 128   # KRYLOV_MAXCOL of csrc/arnoldi.cuh


class LegolasError(RuntimeError):
    """logger%error in the reference (src/dataIO/mod_logging.f08:81-85) raises; so do we."""


@dataclass
class SolverSettings:
    """solver_settings_t (src/settings/mod_solver_settings.f08:30-44)."""

    solver: str = "QR-invert"
    arpack_mode: str = "general"
    number_of_eigenvalues: int = 10
    which_eigenvalues: str = "LM"
    maxiter: int = 0
    ncv: int = 0
    sigma: complex = complex(float("nan"), float("nan"))
    tolerance: float = 5.0e-15
    refine_steps: int = 0   # legolas_b200 extension: iterative-refinement sweeps per solve


@dataclass
class Settings:
    """The part of settings_t (+ mod_equilibrium_params k2, k3) the hot path reads."""

    gridpts: int = 51
    geometry: str = "Cartesian"
    physics_type: str = "mhd"
    k2: float = 0.0
    k3: float = 0.0
    gamma: float = 5.0 / 3.0
    incompressible: bool = False
    flow: bool = False
    resistivity: bool = False
    cooling: bool = False
    heating: bool = False
    conduction: bool = False
    perpendicular_conduction: bool = False
    parallel_conduction: Optional[bool] = None   # None: on whenever conduction is (datfile header only)
    viscosity: bool = False
    viscosity_value: float = 0.0
    viscous_heating: bool = False
    hall: bool = False
    electron_inertia: bool = False
    electron_fraction: float = 0.5
    gravity: bool = False
    boundary_type: str = "wall"
    coaxial: bool = False
    gauss_nodes: Optional[np.ndarray] = None     # None = Legolas 2.0.6 constants
    gauss_weights: Optional[np.ndarray] = None
    solvers: SolverSettings = field(default_factory=SolverSettings)

    @property
    def nb_eqs(self) -> int:
        return {"mhd": 8, "hd": 5, "hd-1d": 3}[self.physics_type]

    @property
    def dim_matrix(self) -> int:     # src/settings/mod_dims.f08:36-44
        return self.gridpts * 2 * self.nb_eqs

    def to_c(self) -> CSettings:
        if self.geometry not in GEOMETRIES:
            raise LegolasError("geometry has no defined scale factor")
        if self.boundary_type not in BOUNDARY_TYPES:
            raise LegolasError(f"unknown boundary_type {self.boundary_type}")
        if self.physics_type not in PHYSICS_TYPES:
            raise LegolasError(f"unknown physics_type {self.physics_type}")
        cs = CSettings()
        cs.gridpts = self.gridpts
        cs.physics_type = PHYSICS_TYPES[self.physics_type]
        cs.geometry = GEOMETRIES[self.geometry]
        for key in ("incompressible", "flow", "resistivity", "cooling", "heating", "conduction",
                    "perpendicular_conduction", "viscosity", "viscous_heating", "hall",
                    "electron_inertia", "gravity", "coaxial"):
            setattr(cs, key, int(bool(getattr(self, key))))
        cs.boundary_type = BOUNDARY_TYPES[self.boundary_type]
        cs.k2, cs.k3, cs.gamma = self.k2, self.k3, self.gamma
        cs.viscosity_value = self.viscosity_value
        cs.electron_fraction = self.electron_fraction
        if self.gauss_nodes is not None:
            for i in range(4):
                cs.gauss_nodes[i] = float(self.gauss_nodes[i])
                cs.gauss_weights[i] = float(self.gauss_weights[i])
        return cs


@dataclass
class ArpackConfig:
    """arpack_t after new_arpack_config."""

    evpdim: int
    mode: int
    bmat: str
    nev: int
    ncv: int
    maxiter: int
    which: str
    tolerance: float
    residual: np.ndarray
    info: int = 1            # start vector supplied by us (mod_arpack_type.f08:210)
    iparam: Dict[int, int] = field(default_factory=dict)


def zlarnv(n: int, seed=ZLARNV_SEED) -> np.ndarray:
    """LAPACK zlarnv(idist=2, iseed) — bit-exact (host utility of the library)."""
    lib = _lib.load()
    iseed = (C.c_int32 * 4)(*seed)
    out = np.empty(n, dtype=np.complex128)
    rc = lib.lgpu_zlarnv(iseed, n, out.ctypes.data)
    if rc != 0:
        raise LgpuError(rc, "lgpu_zlarnv")
    return out


def new_arpack_config(evpdim: int, mode: int, bmat: str, solver_settings: SolverSettings) -> ArpackConfig:
    """Validation + defaults of mod_arpack_type.f08:74-275; writes the effective ncv /
    maxiter back into ``solver_settings`` like the reference (:96-99)."""
    if mode not in (1, 2, 3):
        raise LegolasError(f"Arnoldi: mode = {mode} is invalid, expected 1, 2 or 3")
    if bmat not in ("I", "G"):
        raise LegolasError(f"Arnoldi: bmat = {bmat} is invalid, expected 'I' or 'G'")
    which = solver_settings.which_eigenvalues
    if which not in ALLOWED_WHICH:
        raise LegolasError(f"Arnoldi: which_eigenvalues = {which} is invalid, expected one of "
                           f"{list(ALLOWED_WHICH)}")
    nev = solver_settings.number_of_eigenvalues
    if nev <= 0:
        raise LegolasError(f"Arnoldi: number of eigenvalues must be >= 0 but got {nev}")
    if nev >= evpdim:
        raise LegolasError(f"Arnoldi: number of eigenvalues ({nev}) >= matrix size ({evpdim})")
    ncv = solver_settings.ncv
    if ncv == 0:
        ncv = max(nev + 1, min(2 * nev, evpdim))
    if ncv - nev < 1:
        raise LegolasError(f"ncv too low, expected ncv - nev >= 1 but got ncv - nev = {ncv - nev}")
    if ncv > evpdim:
        raise LegolasError(f"ncv too high, expected ncv < N but got ncv = {ncv} and N = {evpdim}")
    if ncv > MAX_NCV:
        raise LegolasError(f"ncv = {ncv} exceeds the {MAX_NCV} basis columns the device Arnoldi kernels hold "
                           f"(number_of_eigenvalues <= {MAX_NCV // 2} with the default ncv = 2 nev)")
    maxiter = solver_settings.maxiter
    if maxiter < 0:
        raise LegolasError(f"Arnoldi: maxiter must be positive, but is equal to {maxiter}")
    if maxiter == 0:
        maxiter = max(100, 10 * nev)
    solver_settings.maxiter = maxiter
    solver_settings.ncv = ncv
    cfg = ArpackConfig(evpdim=evpdim, mode=mode, bmat=bmat, nev=nev, ncv=ncv, maxiter=maxiter,
                       which=which, tolerance=solver_settings.tolerance, residual=zlarnv(evpdim))
    cfg.iparam = {1: 1, 3: maxiter, 7: mode}
    return cfg


class Context:
    """Owns the device-resident A, B, the BCR factors and the Krylov basis (lgpu_ctx)."""

    def __init__(self, device: int = 0, log_level: int = 0):
        self._lib = _lib.load()
        handle = C.c_void_p()
        rc = self._lib.lgpu_create(C.byref(handle), device, log_level)
        if rc != 0:
            raise LgpuError(rc, "lgpu_create failed (no CUDA device? there is no CPU fallback)")
        self._h = handle
        self.device = device
        self._keepalive = None
        # 2 * nb_eqs and the state vector of the resident matrices (mod_settings.f08:69-86)
        self.dim_subblock = 16
        self.state_vector = STATE_VECTORS["mhd"]

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self._lib.lgpu_last_error(self._h)
            raise LgpuError(rc, f"{what}: {msg.decode() if msg else ''}")

    # ---- plumbing
    def set_stream(self, cuda_stream: int):
        self._check(self._lib.lgpu_set_stream(self._h, C.c_void_p(cuda_stream)), "set_stream")

    def synchronize(self):
        self._check(self._lib.lgpu_synchronize(self._h), "synchronize")

    def set_sm_limit(self, max_sms: int) -> None:
        """Cap the CTAs of this context's fused Gram-Schmidt step (device-wide barrier) so that several contexts of
        one process can have calls in flight on one GPU (lgpu_set_sm_limit); 0 = the whole device."""
        self._check(self._lib.lgpu_set_sm_limit(self._h, int(max_sms)), "set_sm_limit")

    @property
    def dim(self) -> int:
        n = C.c_int32()
        self._check(self._lib.lgpu_matrix_dim(self._h, C.byref(n)), "matrix_dim")
        return n.value

    def counters(self, reset: bool = False) -> int:
        n = C.c_int64()
        self._check(self._lib.lgpu_counters(self._h, C.byref(n), int(reset)), "counters")
        return n.value

    def phase_times(self) -> dict:
        t = [C.c_double() for _ in range(4)]
        self._check(self._lib.lgpu_phase_times(self._h, *[C.byref(x) for x in t]), "phase_times")
        return dict(zip(("assemble_ms", "factor_ms", "iter_ms", "extract_ms"), (x.value for x in t)))

    KINDS = ("assemble", "factor", "matvec", "fwd_stage0", "fwd_stage", "top_stage", "bwd_stage",
             "bwd_stage0", "dots", "update", "scale", "gemm", "cgs2_step", "other")

    def set_profiling(self, enable: bool):
        self._check(self._lib.lgpu_set_profiling(self._h, int(enable)), "set_profiling")

    def profile(self, reset: bool = False) -> dict:
        """{kernel class: (total ms, launches, algorithmic bytes)}; the times are CUDA-event
        durations measured on the launching stream around every launch of that class."""
        n = len(self.KINDS)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        nbytes = (C.c_double * n)()
        self._check(self._lib.lgpu_profile_read(self._h, ms, cnt, nbytes, n, int(reset)), "profile_read")
        return {k: (ms[i], cnt[i], nbytes[i]) for i, k in enumerate(self.KINDS)}

    # ---- assembly
    def assemble(self, settings: Settings, grid: np.ndarray, gauss_grid: np.ndarray,
                 fields: Dict[str, np.ndarray]):
        """Host arrays in, A and B resident on the device out."""
        unknown = set(fields) - set(FIELD_NAMES)
        if unknown:
            raise LegolasError(f"unknown field(s) {sorted(unknown)}")
        grid = np.ascontiguousarray(grid, dtype=np.float64)
        gauss_grid = np.ascontiguousarray(gauss_grid, dtype=np.float64)
        if grid.shape != (settings.gridpts,) or gauss_grid.shape != (4 * (settings.gridpts - 1),):
            raise LegolasError("grid / gaussian grid size does not match settings%grid")
        arrs, ptrs = [], (C.c_void_p * N_FIELDS)()
        for i, name in enumerate(FIELD_NAMES):
            val = fields.get(name)
            if val is None:
                ptrs[i] = None
                continue
            a = np.ascontiguousarray(np.broadcast_to(np.asarray(val, dtype=np.float64),
                                                     gauss_grid.shape))
            arrs.append(a)
            ptrs[i] = a.ctypes.data
        cs = settings.to_c()
        self._check(self._lib.lgpu_assemble(self._h, C.byref(cs), grid.ctypes.data,
                                            gauss_grid.ctypes.data, ptrs), "assemble")
        self._resident(settings)

    def _resident(self, settings: Optional[Settings]):
        ptype = settings.physics_type if settings is not None else "mhd"
        self.state_vector = STATE_VECTORS[ptype]
        self.dim_subblock = 2 * len(self.state_vector)

    def assemble_device(self, settings: Settings, grid_ptr: int, gauss_ptr: int, field_ptrs):
        """Same with device pointers (inputs already resident in HBM)."""
        ptrs = (C.c_void_p * N_FIELDS)()
        for i in range(N_FIELDS):
            ptrs[i] = field_ptrs[i] if field_ptrs[i] else None
        cs = settings.to_c()
        self._check(self._lib.lgpu_assemble_device(self._h, C.byref(cs), C.c_void_p(grid_ptr),
                                                   C.c_void_p(gauss_ptr), ptrs), "assemble_device")
        self._resident(settings)

    def export_blocks(self, which: str) -> np.ndarray:
        """(gridpts, 3, d, d) complex, d = dim_subblock (16 mhd / 10 hd / 6 hd-1d): [sub, diag, super]
        blocks in the reference's numbering; [b, t, i, j] = row i, col j."""
        n = self.dim
        d = self.dim_subblock
        g = n // d
        raw = np.empty((g, 3, d, d), dtype=np.complex128)   # device blocks are column-major
        self._check(self._lib.lgpu_export_blocks(self._h, _which(which), raw.ctypes.data),
                    "export_blocks")
        return np.ascontiguousarray(raw.transpose(0, 1, 3, 2))

    def export_coo(self, which: str):
        """Triplets (rows, cols, vals), 1-based, in the reference's datfile order."""
        nnz = C.c_int64()
        w = _which(which)
        self._check(self._lib.lgpu_export_coo(self._h, w, C.byref(nnz), None, None, None), "export_coo")
        rows = np.empty(nnz.value, dtype=np.int32)
        cols = np.empty(nnz.value, dtype=np.int32)
        vals = np.empty(nnz.value, dtype=np.complex128)
        self._check(self._lib.lgpu_export_coo(self._h, w, C.byref(nnz), rows.ctypes.data,
                                              cols.ctypes.data, vals.ctypes.data), "export_coo")
        return rows, cols, vals

    def import_coo(self, which: str, n: int, rows, cols, vals):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        vals = np.ascontiguousarray(vals, dtype=np.complex128)
        self._check(self._lib.lgpu_import_coo(self._h, _which(which), n, len(rows), rows.ctypes.data,
                                              cols.ctypes.data, vals.ctypes.data), "import_coo")
        self._resident(None)

    # ---- linear algebra pieces
    def factorize(self, sigma: complex) -> int:
        info = C.c_int32()
        self._check(self._lib.lgpu_factorize(self._h, sigma.real, sigma.imag, C.byref(info)), "factorize")
        return info.value

    def _vec(self, x, what: str) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.complex128)
        if x.shape != (self.dim,):
            raise LegolasError(f"{what}: vector of shape {x.shape}, matrix dimension is {self.dim}")
        return x

    def solve(self, rhs, refine_steps: int = 0) -> np.ndarray:
        rhs = self._vec(rhs, "solve")
        x = np.empty_like(rhs)
        self._check(self._lib.lgpu_solve(self._h, rhs.ctypes.data, x.ctypes.data, refine_steps), "solve")
        return x

    def matvec(self, which: str, x) -> np.ndarray:
        x = self._vec(x, "matvec")
        y = np.empty_like(x)
        self._check(self._lib.lgpu_matvec(self._h, _which(which), x.ctypes.data, y.ctypes.data), "matvec")
        return y

    def apply_op(self, x, refine_steps: int = 0) -> np.ndarray:
        x = self._vec(x, "apply_op")
        y = np.empty_like(x)
        self._check(self._lib.lgpu_apply_op(self._h, x.ctypes.data, y.ctypes.data, refine_steps), "apply_op")
        return y

    def apply_op_device(self, x_ptr: int, y_ptr: int, repeat: int = 1, refine_steps: int = 0) -> float:
        """``repeat`` operator applications y = (A - sigma B)^-1 B x on device vectors (raw pointers), back to back;
        returns the CUDA-event time per application in milliseconds (lgpu_apply_op_device)."""
        ms = C.c_double(0.0)
        self._check(self._lib.lgpu_apply_op_device(self._h, C.c_void_p(x_ptr), C.c_void_p(y_ptr), refine_steps, repeat,
                                                   C.byref(ms)), "apply_op_device")
        return ms.value

    # ---- the eigen-solve
    def arnoldi_general(self, cfg: ArpackConfig, refine_steps: int = 0, want_vectors: bool = True):
        """solve_arpack_general (src/solvers/arnoldi/smod_arpack_general.f08:14-131): ARPACK on
        OP = B^-1 A (mode 1, bmat "I"); returns (omega, vr, stats) like ``shift_invert``."""
        return self.shift_invert(cfg, 0j, refine_steps, want_vectors, _entry="lgpu_arnoldi_general")

    def shift_invert(self, cfg: ArpackConfig, sigma: complex, refine_steps: int = 0,
                     want_vectors: bool = True, vr_view: bool = False, _entry: str = "lgpu_shift_invert"):
        """``vr_view=True`` returns the Ritz vectors as a view of the context's page-locked
        read-back buffer (valid until the next solve on this context) instead of a copy."""
        ca = _arnoldi_c(cfg, sigma, refine_steps)
        resid = np.ascontiguousarray(cfg.residual, dtype=np.complex128)
        if cfg.evpdim != self.dim or resid.shape != (self.dim,):
            raise LegolasError(f"Arnoldi: evpdim = {cfg.evpdim} / start vector of shape {resid.shape} do not match "
                               f"the resident matrices (dimension {self.dim})")
        omega = np.empty(cfg.nev, dtype=np.complex128)
        vr = None
        if want_vectors:
            nbytes = cfg.evpdim * cfg.nev * 16
            # reuse one page-locked buffer per context for large read-backs
            if nbytes >= (1 << 20):
                if getattr(self, "_vr_pinned", None) is None or self._vr_pinned.shape != (cfg.evpdim, cfg.nev):
                    self._vr_pinned = pinned_empty((cfg.evpdim, cfg.nev))
                vr = self._vr_pinned
            else:
                vr_view = True
                vr = np.empty((cfg.evpdim, cfg.nev), dtype=np.complex128, order="F")
        st = CStats()
        self._check(getattr(self._lib, _entry)(
            self._h, C.byref(ca), resid.ctypes.data, omega.ctypes.data,
            vr.ctypes.data if vr is not None else None, C.byref(st)), _entry[5:])
        if vr is not None and not vr_view:
            vr = np.array(vr, order="F")
        return omega, vr, _stats_dict(st)

    def residuals(self, omega, vr) -> np.ndarray:
        """get_residual for every pair (src/dataIO/mod_output.f08:511-545)."""
        omega = np.ascontiguousarray(omega, dtype=np.complex128)
        vr = np.asfortranarray(vr, dtype=np.complex128)
        if vr.shape != (self.dim, len(omega)):
            raise LegolasError(f"eigenvector array has shape {vr.shape}, expected {(self.dim, len(omega))}")
        res = np.zeros(len(omega), dtype=np.float64)
        self._check(self._lib.lgpu_residuals(self._h, len(omega), omega.ctypes.data, vr.ctypes.data,
                                             res.ctypes.data_as(C.POINTER(C.c_double))), "residuals")
        return res

    def eigenfunctions(self, vr, idxs, state_vector=None) -> dict:
        """base_ef_t%assemble for every variable of the state vector (src/eigenfunctions/mod_base_efs.f08:35-61):
        {name: (2 G - 1, len(idxs)) complex}; ``idxs`` are 1-based columns of ``vr`` (idxs_to_assemble)."""
        vr = np.asfortranarray(vr, dtype=np.complex128)
        idxs = np.ascontiguousarray(idxs, dtype=np.int32)
        if vr.shape[0] != self.dim or (len(idxs) and (idxs.min() < 1 or idxs.max() > vr.shape[1])):
            raise LegolasError("eigenfunctions: eigenvector array / indices do not match the matrices")
        state_vector = state_vector or self.state_vector
        npts = 2 * (self.dim // self.dim_subblock) - 1
        out = np.zeros((self.dim_subblock // 2, len(idxs), npts), dtype=np.complex128)
        self._check(self._lib.lgpu_eigenfunctions(self._h, vr.ctypes.data, len(idxs),
                                                  idxs.ctypes.data_as(C.POINTER(C.c_int32)), out.ctypes.data),
                    "eigenfunctions")
        return {name: out[p].T.copy(order="F") for p, name in enumerate(state_vector)}

    def inverse_iteration(self, sigma: complex, maxiter: int = 0, tolerance: float = 5.0e-15,
                          want_vector: bool = True):
        """inverse_iteration (src/solvers/smod_inverse_iteration.f08:16-205): (omega, x, stats)."""
        omega = np.zeros(1, dtype=np.complex128)
        x = np.empty(self.dim, dtype=np.complex128) if want_vector else None
        st = CStats()
        self._check(self._lib.lgpu_inverse_iteration(self._h, sigma.real, sigma.imag, int(maxiter), float(tolerance),
                                                     omega.ctypes.data, x.ctypes.data if want_vector else None,
                                                     C.byref(st)), "inverse_iteration")
        return complex(omega[0]), x, _stats_dict(st)

    def shift_invert_device(self, cfg: ArpackConfig, sigma: complex, resid_ptr: int, vr_ptr: int,
                            refine_steps: int = 0):
        ca = _arnoldi_c(cfg, sigma, refine_steps)
        omega = np.empty(cfg.nev, dtype=np.complex128)
        st = CStats()
        self._check(self._lib.lgpu_shift_invert_device(
            self._h, C.byref(ca), C.c_void_p(resid_ptr), omega.ctypes.data,
            C.c_void_p(vr_ptr) if vr_ptr else None, C.byref(st)), "shift_invert_device")
        return omega, _stats_dict(st)


class _PinnedBlock:
    """Owner of one page-locked host allocation (freed when the last array view dies)."""

    def __init__(self, lib, nbytes: int):
        self._lib = lib
        self.ptr = lib.lgpu_host_alloc(nbytes)
        if not self.ptr:
            raise MemoryError("lgpu_host_alloc failed")
        self.nbytes = nbytes

    def __del__(self):
        try:
            self._lib.lgpu_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype=np.complex128, order="F") -> np.ndarray:
    """numpy array backed by page-locked memory (fast device -> host copies)."""
    lib = _lib.load()
    dt = np.dtype(dtype)
    count = int(np.prod(shape))
    block = _PinnedBlock(lib, max(count * dt.itemsize, 1))
    buf = (C.c_char * block.nbytes).from_address(block.ptr)
    arr = np.frombuffer(buf, dtype=dt, count=count).reshape(shape, order=order)
    buf._owner = block          # the ctypes buffer (base of the array) keeps the allocation alive
    return arr


def _which(which: str) -> int:
    if which not in ("A", "B"):
        raise LegolasError(f"invalid or empty matrix label: {which}")
    return 0 if which == "A" else 1


def _arnoldi_c(cfg: ArpackConfig, sigma: complex, refine_steps: int) -> CArnoldi:
    ca = CArnoldi()
    ca.nev, ca.ncv, ca.maxiter = cfg.nev, cfg.ncv, cfg.maxiter
    ca.which = cfg.which.encode()
    ca.tol = cfg.tolerance
    ca.sigma_re, ca.sigma_im = sigma.real, sigma.imag
    ca.refine_steps = refine_steps
    return ca


def _stats_dict(st: CStats) -> dict:
    return {name: getattr(st, name) for name, _ in CStats._fields_ if name != "reserved"}


def _parse_arnoldi_status(stats: dict, cfg) -> None:
    """parse_znaupd_info / parse_zneupd_info (src/solvers/arnoldi/mod_arpack_type.f08:364-436) and
    the zgbtrf info check (src/solvers/mod_linear_systems.f08:125-137): info = 1 (maxiter reached)
    and a singular pivot are warnings, every other non-zero info is logger%error, i.e. raises."""
    import warnings

    if stats.get("lu_info", 0) != 0:
        warnings.warn(f"LAPACK routine zgbtrf failed! info = {stats['lu_info']}", RuntimeWarning, stacklevel=3)
    info = stats["info"]
    if info == 0:
        return
    if info == 1:
        warnings.warn(f"ARPACK failed to converge! (maxiter reached) number of iterations: {cfg.maxiter}, "
                      f"number of converged eigenvalues: {stats['nconv']} / {cfg.nev}", RuntimeWarning, stacklevel=3)
        return
    messages = {3: "znaupd: no shifts could be applied during a cycle of the Arnoldi iteration. Try increasing "
                   "the size of ncv relative to number_of_eigenvalues.",
                -8: "znaupd: error LAPACK eigenvalue calculation",
                -9: "znaupd: starting vector is zero, try rerunning?"}
    raise LegolasError(messages.get(info, f"znaupd: unexpected info = {info} encountered"))


# ------------------------------------------------------------------ reference-named entry points
@dataclass
class Matrices:
    """Stand-in for the (matrix_A, matrix_B) pair of matrix_t: device-resident."""

    ctx: Context
    settings: Settings


def build_matrices(settings: Settings, grid: np.ndarray, gauss_grid: np.ndarray,
                   fields: Dict[str, np.ndarray], ctx: Optional[Context] = None) -> Matrices:
    """``call build_matrices(matrix_B, matrix_A, settings, grid, background, physics)``.

    ``fields`` are the background / physics procedure pointers sampled at ``grid%gaussian_grid``
    (the Fortran shim samples them with ``from_function``); see ``FIELD_NAMES`` for the slots."""
    ctx = ctx or Context()
    ctx.assemble(settings, grid, gauss_grid, fields)
    return Matrices(ctx=ctx, settings=settings)


def solve_evp(matrices: Matrices, settings: Settings, vr_view: bool = False, want_vectors: bool = True):
    """``call solve_evp(matrix_A, matrix_B, settings, omega, right_eigenvectors)`` for
    ``solver = "arnoldi"``, ``arpack_mode = "shift-invert"`` or ``"general"``.  Returns (omega, vr, arpack_cfg, stats);
    omega(nconv:) is NaN when ARPACK-style convergence was not reached for all nev (a warning,
    not an error, in the reference: mod_arpack_type.f08:375-381).  ``want_vectors = False`` is the reference's
    ``settings%io%should_compute_eigenvectors() == .false.`` (src/main.f08:124-153: ``right_eigenvectors`` is then a
    dummy of size (2, 2) and zneupd is asked for eigenvalues only): vr is returned as None."""
    sv = settings.solvers
    if sv.solver == "inverse-iteration":
        # mod_solvers.f08:92-123 dispatch; one eigenpair, omega(1) / vr(:, 1)
        sigma = complex(sv.sigma)
        if sigma != sigma:
            raise LegolasError("sigma must be set for the inverse-iteration solver")
        if sv.maxiter < 0:
            raise LegolasError(f"maxiter has to be positive, but is equal to {sv.maxiter}")
        if sigma == 0:
            raise LegolasError("inverse-iteration: sigma can not be equal to zero")
        omega, x, stats = matrices.ctx.inverse_iteration(sigma, sv.maxiter, sv.tolerance)
        return np.array([omega]), x.reshape(-1, 1), None, stats
    if sv.solver != "arnoldi":
        raise LegolasError(f"solver {sv.solver!r} stays on the Fortran host; 'arnoldi' and "
                           "'inverse-iteration' are built here")
    if sv.arpack_mode == "general":
        # smod_arpack_main.f08:57-65: mode = 1, bmat = "I"; OP = B^-1 A
        cfg = new_arpack_config(matrices.ctx.dim, mode=1, bmat="I", solver_settings=sv)
        omega, vr, stats = matrices.ctx.arnoldi_general(cfg, sv.refine_steps)
        cfg.info = stats["info"]
        cfg.iparam.update({5: stats["nconv"], 9: stats["n_op"], 10: stats["n_bx"], 11: stats["n_reorth"]})
        _parse_arnoldi_status(stats, cfg)
        return omega, vr, cfg, stats
    if sv.arpack_mode != "shift-invert":
        # smod_arpack_main.f08:78-82
        raise LegolasError(f"unknown mode for ARPACK: {sv.arpack_mode}")
    if math.isnan(sv.sigma.real) or math.isnan(sv.sigma.imag):
        raise LegolasError("sigma is not set")
    cfg = new_arpack_config(matrices.ctx.dim, mode=2, bmat="I", solver_settings=sv)
    omega, vr, stats = matrices.ctx.shift_invert(cfg, complex(sv.sigma), sv.refine_steps, vr_view=vr_view,
                                                 want_vectors=want_vectors)
    cfg.info = stats["info"]
    cfg.iparam.update({5: stats["nconv"], 9: stats["n_op"], 10: stats["n_bx"], 11: stats["n_reorth"]})
    _parse_arnoldi_status(stats, cfg)
    return omega, vr, cfg, stats
