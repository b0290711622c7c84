"""Host-side grid, background and physics sampling (the part that stays on the host).

In the reference the equilibrium is a Fortran plug-in that registers ``real(dp)`` procedure
pointers in ``background_t`` / ``physics_t`` (src/mod_equilibrium.f08:224-314,
src/background/mod_background.f08:44-131); the C-ABI shim samples them at
``grid%gaussian_grid`` before calling the GPU.  This module does the same for the benchmark
configurations so that ``bench.py`` and the examples have inputs without a Fortran host:
a ``Background`` holds callables, ``sample`` evaluates them on the Gaussian grid.

  grid ................... src/mod_grid.f08:121-140,160-196, src/settings/mod_grid_settings.f08:112-138
  adiabatic_homo ......... src/equilibria/smod_equil_adiabatic_homo.f08:21-54
  suydam_cluster ......... src/equilibria/smod_equil_suydam_cluster.f08:28-120
  resistive_tearing ...... src/equilibria/smod_equil_resistive_tearing.f08:26-89
  magnetothermal ......... src/equilibria/smod_equil_magnetothermal_instabilities.f08:34-86
  kelvin_helmholtz_cd .... src/equilibria/smod_equil_kelvin_helmholtz_cd.f08:32-98
  MRI_accretion .......... src/equilibria/smod_equil_MRI_accretion.f08:36-133
  couette_flow ........... src/equilibria/smod_equil_couette_flow.f08:25-82
  units / physics ........ src/settings/mod_units.f08:161-211, src/physics/mod_thermal_conduction.f08:56-247,
                           src/physics/mod_heatloss.f08:54-139, src/physics/cooling_curves/*
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Tuple

import numpy as np
from scipy import special

from .api import Settings

GAUSSIAN_NODES = (-0.861136311594053, -0.339981043584856, 0.339981043584856, 0.861136311594053)


# ----------------------------------------------------------------------------- grid_t
class Grid:
    def __init__(self, settings: Settings, grid_start: float, grid_end: float,
                 force_r0: bool = False, nodes=GAUSSIAN_NODES):
        if settings.geometry == "cylindrical" and not force_r0 and abs(grid_start) <= 5e-15:
            grid_start = 0.025     # avoid the on-axis singularity (mod_grid_settings.f08:132-138)
        if grid_start > grid_end:
            raise ValueError(f"grid generation: grid start = {grid_start} > grid end = {grid_end}")
        pts = settings.gridpts
        step = (grid_end - grid_start) / (pts - 1)
        xbar = [grid_start]
        for _ in range(pts - 1):
            xbar.append(xbar[-1] + step)
        kappa = (grid_end - xbar[pts - 2]) / (xbar[pts - 1] - xbar[pts - 2])
        base = np.empty(pts)
        base[0] = grid_start
        base[1:] = np.asarray(xbar[:-1]) + kappa * step
        self.base_grid = base
        lo, hi = base[:-1, None], base[1:, None]
        self.gaussian_grid = (0.5 * (hi - lo) * np.asarray(nodes)[None, :] + 0.5 * (lo + hi)).ravel()


# ----------------------------------------------------------------------- background_t
Func = Callable[[np.ndarray], np.ndarray]


class Background:
    """Named slots of real functions of position, default ``zero_func``."""

    def __init__(self):
        self.funcs: Dict[str, Func] = {}

    def set(self, **funcs: Func) -> "Background":
        self.funcs.update(funcs)
        return self

    def sample(self, x: np.ndarray) -> Dict[str, np.ndarray]:
        out = {}
        for name, fn in self.funcs.items():
            out[name] = np.broadcast_to(np.asarray(fn(x), dtype=np.float64), x.shape).copy()
        return out


def _const(value: float) -> Func:
    return lambda x: np.full_like(x, value)


# ------------------------------------------------------------------------------ units
class UnitSystem:
    """cgs unit system derived from (length, magnetic field, temperature)."""

    MP, KB, MU0 = 1.672621777e-24, 1.3806488e-16, 4.0 * math.pi

    def __init__(self, unit_length, unit_magneticfield, unit_temperature, mean_molecular_weight):
        self.unit_length = unit_length
        self.unit_magneticfield = unit_magneticfield
        self.unit_temperature = unit_temperature
        self.unit_pressure = unit_magneticfield ** 2 / self.MU0
        self.unit_density = (mean_molecular_weight * self.unit_pressure * self.MP
                             / (self.KB * unit_temperature))
        self.unit_numberdensity = self.unit_density / self.MP
        self.unit_velocity = unit_magneticfield / math.sqrt(self.MU0 * self.unit_density)
        self.unit_time = unit_length / self.unit_velocity
        self.unit_lambdaT = self.unit_pressure / (self.unit_time * self.unit_numberdensity ** 2)
        self.unit_conduction = (self.unit_density * unit_length * self.unit_velocity ** 3
                                / unit_temperature)


_ROSNER_LOGT = (3.89063, 4.30195, 4.575, 4.9, 5.4, 5.77, 6.315, 7.60457)
_ROSNER_LOGXI = (-69.900, -48.307, -21.850, -31.000, -21.200, -10.400, -21.940, -17.730, -26.602)
_ROSNER_ALPHA = (11.7, 6.15, 0.0, 2.0, 0.0, -2.0, 0.0, -0.666666667, 0.5)


def _rosner_piece(log_t: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    idx = np.searchsorted(np.asarray(_ROSNER_LOGT), log_t, side="right")
    # the reference picks the first j with logT < logT_j, and piece 9 above the table
    return np.asarray(_ROSNER_LOGXI)[idx], np.asarray(_ROSNER_ALPHA)[idx]


def rosner_cooling(T0: np.ndarray, units: UnitSystem) -> Tuple[np.ndarray, np.ndarray]:
    """(lambda(T), dlambda/dT) in code units for the piecewise Rosner curve."""
    log_t = np.log10(T0 * units.unit_temperature)
    logxi, alpha = _rosner_piece(log_t)
    lam = 10.0 ** (logxi + alpha * log_t) / units.unit_lambdaT
    dlam = (alpha * 10.0 ** (logxi + (alpha - 1.0) * log_t)) / (units.unit_lambdaT / units.unit_temperature)
    return lam, dlam


def spitzer_parallel_conduction(T0: np.ndarray, units: UnitSystem) -> Tuple[np.ndarray, np.ndarray]:
    pf, coulomb_log = 1.8e-5, 22.0
    T = T0 * units.unit_temperature
    kappa = (pf * T ** 2.5 / coulomb_log) / units.unit_conduction
    dkappa = (pf * 2.5 * T ** 1.5 / coulomb_log) / (units.unit_conduction / units.unit_temperature)
    return kappa, dkappa


# ------------------------------------------------------------------------- equilibria
def adiabatic_homo(gridpts: int, k2=0.0, k3=math.pi, rho0=1.0, T0=1.0, B02=0.0, B03=1.0):
    s = Settings(gridpts=gridpts, geometry="Cartesian", k2=k2, k3=k3)
    grid = Grid(s, 0.0, 1.0)
    bg = Background().set(rho0=_const(rho0), T0=_const(T0), B02=_const(B02), B03=_const(B03))
    return s, grid, bg.sample(grid.gaussian_grid)


def couette_flow(gridpts: int, k2=0.0, k3=1.0, rho0=1.0, T0=1.0, v02=0.0, v03=1.0, viscosity_value=1.0e-3,
                 physics_type="hd"):
    """Plane Couette flow between two walls, flow + viscosity; runs as "hd" in the reference's
    regression suite (tests/regression_tests/test_couette_flow_HD.py:24-50)."""
    s = Settings(gridpts=gridpts, geometry="Cartesian", k2=k2, k3=k3, physics_type=physics_type, flow=True,
                 viscosity=True, viscosity_value=viscosity_value)
    grid = Grid(s, 0.0, 1.0)
    width = 1.0
    bg = Background().set(rho0=_const(rho0), T0=_const(T0),
                          v02=lambda x: v02 * x / width, dv02=_const(v02 / width),
                          v03=lambda x: v03 * x / width, dv03=_const(v03 / width))
    return s, grid, bg.sample(grid.gaussian_grid)


def suydam_cluster(gridpts: int, k2=1.0, k3=-1.2, rho0=1.0, v02=0.0, v03=0.14, p0=0.05, p1=0.1,
                   alpha=2.0):
    s = Settings(gridpts=gridpts, geometry="cylindrical", flow=True, k2=k2, k3=k3)
    grid = Grid(s, 0.0, 1.0)
    j0 = lambda r: special.jv(0, alpha * r)
    j1 = lambda r: special.jv(1, alpha * r)
    j2 = lambda r: special.jv(2, alpha * r)
    root = math.sqrt(1.0 - p1)
    bg = Background().set(
        rho0=_const(rho0),
        T0=lambda r: (p0 + 0.5 * p1 * j0(r) ** 2) / rho0,
        dT0=lambda r: p1 * j0(r) * (-alpha * j1(r)) / rho0,
        v02=_const(v02),
        v03=lambda r: v03 * (1.0 - r ** 2),
        dv03=lambda r: -2.0 * v03 * r,
        B02=j1,
        dB02=lambda r: alpha * (0.5 * j0(r) - 0.5 * j2(r)),
        B03=lambda r: root * j0(r),
        dB03=lambda r: -alpha * root * j1(r),
    )
    return s, grid, bg.sample(grid.gaussian_grid)


def resistive_tearing(gridpts: int, k2=0.49, k3=0.0, alpha=4.73884, beta=0.15, rho0=1.0, eta=1.0e-4):
    s = Settings(gridpts=gridpts, geometry="Cartesian", resistivity=True, k2=k2, k3=k3)
    grid = Grid(s, -0.5, 0.5)
    bg = Background().set(
        rho0=_const(rho0),
        T0=lambda x: beta * np.sqrt(np.sin(alpha * x) ** 2 + np.cos(alpha * x) ** 2) / 2.0,
        B02=lambda x: np.sin(alpha * x),
        dB02=lambda x: alpha * np.cos(alpha * x),
        ddB02=lambda x: -alpha ** 2 * np.sin(alpha * x),
        B03=lambda x: np.cos(alpha * x),
        dB03=lambda x: -alpha * np.sin(alpha * x),
        ddB03=lambda x: -alpha ** 2 * np.cos(alpha * x),
        eta=_const(eta),
    )
    return s, grid, bg.sample(grid.gaussian_grid)


def magnetothermal_instabilities(gridpts: int, k2=0.0, k3=1.0, T0=1.0):
    s = Settings(gridpts=gridpts, geometry="cylindrical", cooling=True, heating=True,
                 conduction=True, perpendicular_conduction=False, k2=k2, k3=k3)
    units = UnitSystem(unit_length=1.0e8, unit_magneticfield=10.0, unit_temperature=2.6e6,
                       mean_molecular_weight=1.0)
    grid = Grid(s, 0.0, 1.0)
    r = grid.gaussian_grid
    bg = Background().set(
        rho0=lambda r: (1.0 / (2.0 * (1.0 + r ** 2) ** 2)) / T0,
        drho0=lambda r: -2.0 * r / (T0 * (r ** 2 + 1.0) ** 3),
        T0=_const(T0),
        B02=lambda r: r / (1.0 + r ** 2),
        dB02=lambda r: (1.0 - r ** 2) / (r ** 4 + 2.0 * r ** 2 + 1.0),
    )
    f = bg.sample(r)
    lam, dlam = rosner_cooling(f["T0"], units)
    kpara, dkpara = spitzer_parallel_conduction(f["T0"], units)
    B0 = np.abs(f["B02"])
    dB0 = f["B02"] * f["dB02"] / B0
    # thermal balance: with dT0 = v01 = B01 = kappa_perp = 0 the enforced heating is rho0*lambda,
    # so L0 = 0; only its derivatives survive (mod_heatloss.f08:54-75,109-139)
    f.update(
        L0=np.zeros_like(r), dLdT=f["rho0"] * dlam, dLdrho=lam,
        tcpara=kpara, dtcparadT=dkpara,
        tcprefactor=kpara / B0 ** 2,
        dtcprefactordr=(-2.0 * kpara * dB0) / B0 ** 3,
    )
    return s, grid, f


def kelvin_helmholtz_cd(gridpts: int, k2=-1.0, k3=None, V=1.63, rho0=1.0, p0=1.0, Bz0=0.25, rc=0.5,
                        rj=1.0):
    Bth0 = 0.4 * (rc ** 2 + rj ** 2) / (rj * rc)
    a = 0.1 * rj
    k3 = math.pi / rj if k3 is None else k3
    s = Settings(gridpts=gridpts, geometry="cylindrical", flow=True, k2=k2, k3=k3)
    grid = Grid(s, 0.0, 2.0 * rj)
    bg = Background().set(
        rho0=_const(rho0),
        T0=lambda r: p0 / rho0 - (Bth0 ** 2 / (2.0 * rho0)) * (1.0 - rc ** 4 / (rc ** 2 + r ** 2) ** 2),
        dT0=lambda r: -(2.0 * Bth0 ** 2 / rho0) * rc ** 4 * r / (r ** 2 + rc ** 2) ** 3,
        v03=lambda r: (V / 2.0) * np.tanh((rj - r) / a),
        dv03=lambda r: -(V / (2.0 * a)) / np.cosh((rj - r) / a) ** 2,
        B02=lambda r: Bth0 * r * rc / (rc ** 2 + r ** 2),
        dB02=lambda r: Bth0 * rc * (rc ** 2 - r ** 2) / (r ** 2 + rc ** 2) ** 2,
        B03=_const(Bz0),
    )
    return s, grid, bg.sample(grid.gaussian_grid)


def mri_accretion(gridpts: int, k2=0.0, k3=70.0, beta=100.0, tau=1.0, nu=0.1, x_start=1.0, x_end=2.0):
    s = Settings(gridpts=gridpts, geometry="cylindrical", flow=True, gravity=True, k2=k2, k3=k3)
    grid = Grid(s, x_start, x_end)
    p1 = nu ** 2
    Bz1 = math.sqrt(2.0 * p1 / (beta * (1.0 + tau ** 2)))
    Bth1 = tau * Bz1
    vth1 = math.sqrt(1.0 - 2.5 * p1 - 0.25 * Bth1 ** 2 - 1.25 * Bz1 ** 2)
    rho = lambda r: r ** -1.5
    drho = lambda r: -1.5 * r ** -2.5
    pres = lambda r: p1 * r ** -2.5
    dpres = lambda r: -2.5 * p1 * r ** -3.5
    bg = Background().set(
        rho0=rho, drho0=drho,
        T0=lambda r: pres(r) / rho(r),
        dT0=lambda r: (dpres(r) * rho(r) - drho(r) * pres(r)) / rho(r) ** 2,
        v02=lambda r: vth1 / np.sqrt(r),
        dv02=lambda r: -0.5 * vth1 * r ** -1.5,
        B02=lambda r: Bth1 * r ** -1.25,
        dB02=lambda r: -1.25 * Bth1 * r ** -2.25,
        B03=lambda r: Bz1 * r ** -1.25,
        dB03=lambda r: -1.25 * Bz1 * r ** -2.25,
        g0=lambda r: 1.0 / r ** 2,
    )
    return s, grid, bg.sample(grid.gaussian_grid)


EQUILIBRIA = {
    "adiabatic_homo": adiabatic_homo,
    "suydam_cluster": suydam_cluster,
    "resistive_tearing": resistive_tearing,
    "magnetothermal_instabilities": magnetothermal_instabilities,
    "kelvin_helmholtz_cd": kelvin_helmholtz_cd,
    "MRI_accretion": mri_accretion,
    "couette_flow": couette_flow,
}
