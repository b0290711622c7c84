"""Restated equilibria + physics closures for the benchmark configs (oracle side).

Test infrastructure (see ``oracle/__init__.py``).  Each ``*_eq`` function returns
``(settings, grid, gauss_grid, fields)`` ready for ``oracle.assembly.build_matrices``,
i.e. what the Fortran host would sample from its procedure pointers at
``grid%gaussian_grid`` before crossing the C ABI.

Follows (reference file:line):
  adiabatic_homo ............ src/equilibria/smod_equil_adiabatic_homo.f08:21-54
  suydam_cluster ............ src/equilibria/smod_equil_suydam_cluster.f08:28-120
  resistive_tearing ......... src/equilibria/smod_equil_resistive_tearing.f08:26-89
  magnetothermal ............ src/equilibria/smod_equil_magnetothermal_instabilities.f08:34-86
  kelvin_helmholtz_cd ....... src/equilibria/smod_equil_kelvin_helmholtz_cd.f08:32-98
  MRI_accretion ............. src/equilibria/smod_equil_MRI_accretion.f08:36-133
  couette_flow .............. src/equilibria/smod_equil_couette_flow.f08:25-82
  on-axis grid shift ........ src/settings/mod_grid_settings.f08:112-138
  units ..................... src/settings/mod_units.f08:161-211, src/mod_physical_constants.f08
  resistivity ............... src/physics/mod_resistivity.f08:49-120
  thermal conduction ........ src/physics/mod_thermal_conduction.f08:56-247
  Rosner cooling ............ src/physics/cooling_curves/mod_cooling_curves.f08:103-162,
                              src/physics/cooling_curves/mod_data_rosner.f08:18-30
  heat loss / balance ....... src/physics/mod_heatloss.f08:54-139
"""
from __future__ import annotations

import numpy as np
from scipy.special import jv

from .assembly import GAUSS_NODES, Settings, base_grid, gaussian_grid

DPI = 3.141592653589793238462643383279
COULOMB_LOG = 22.0
MP_CGS = 1.672621777e-24
KB_CGS = 1.3806488e-16
MU0_CGS = 4.0 * DPI
TC_PF_KAPPA_PARA = 1.8e-5
TC_PF_KAPPA_PERP = 8.2e-13

LOGT_ROSNER = np.array([3.89063, 4.30195, 4.575, 4.9, 5.4, 5.77, 6.315, 7.60457])
LOGXI_ROSNER = np.array([-69.900, -48.307, -21.850, -31.000, -21.200, -10.400,
                         -21.940, -17.730, -26.602])
ALPHA_ROSNER = np.array([11.7, 6.15, 0.0, 2.0, 0.0, -2.0, 0.0, -0.666666667, 0.5])


class Units:
    """units_t set from temperature (mod_units.f08:118-134,161-211)."""

    def __init__(self, unit_length=1.0e9, unit_magneticfield=10.0, unit_temperature=1.0e6,
                 mean_molecular_weight=0.5):
        self.unit_length = unit_length
        self.unit_magneticfield = unit_magneticfield
        self.unit_temperature = unit_temperature
        self.mean_molecular_weight = mean_molecular_weight
        self.unit_pressure = unit_magneticfield**2 / MU0_CGS
        self.unit_density = (
            mean_molecular_weight * self.unit_pressure * MP_CGS / (KB_CGS * unit_temperature)
        )
        self.unit_numberdensity = self.unit_density / MP_CGS
        self.unit_velocity = unit_magneticfield / np.sqrt(MU0_CGS * self.unit_density)
        self.unit_mass = self.unit_density * unit_length**3
        self.unit_time = unit_length / self.unit_velocity
        self.unit_resistivity = unit_length**2 / self.unit_time
        self.unit_lambdaT = self.unit_pressure / (self.unit_time * self.unit_numberdensity**2)
        self.unit_conduction = (
            self.unit_density * unit_length * self.unit_velocity**3 / unit_temperature
        )


def _grid(settings_geometry, start, end, gridpts, nodes, force_r0=False):
    if settings_geometry == "cylindrical" and not force_r0 and abs(start) <= 5e-15:
        start = 0.025
    g = base_grid(start, end, gridpts)
    return g, gaussian_grid(g, nodes)


# ------------------------------------------------------------------------ physics
def rosner_index(logT0):
    """mod_cooling_curves.f08:103-119 (returns 0-based index)."""
    idx = np.empty(logT0.shape, dtype=np.int64)
    for n, val in enumerate(logT0):
        if val > LOGT_ROSNER[7]:
            idx[n] = 8
        else:
            idx[n] = 0
            for j in range(8):
                if val < LOGT_ROSNER[j]:
                    idx[n] = j
                    break
    return idx


def rosner_lambdaT(T0, units: Units):
    logT0 = np.log10(T0 * units.unit_temperature)
    idx = rosner_index(logT0)
    return 10.0 ** (LOGXI_ROSNER[idx] + ALPHA_ROSNER[idx] * logT0) / units.unit_lambdaT


def rosner_dlambdadT(T0, units: Units):
    logT0 = np.log10(T0 * units.unit_temperature)
    idx = rosner_index(logT0)
    alpha = ALPHA_ROSNER[idx]
    return (alpha * 10.0 ** (LOGXI_ROSNER[idx] + (alpha - 1.0) * logT0)) / (
        units.unit_lambdaT / units.unit_temperature
    )


def tcpara(T0, units: Units):
    T = T0 * units.unit_temperature
    return (TC_PF_KAPPA_PARA * T**2.5 / COULOMB_LOG) / units.unit_conduction


def dtcparadT(T0, units: Units):
    T = T0 * units.unit_temperature
    return (TC_PF_KAPPA_PARA * 2.5 * T**1.5 / COULOMB_LOG) / (
        units.unit_conduction / units.unit_temperature
    )


# --------------------------------------------------------------------- equilibria
def adiabatic_homo_eq(gridpts=51, k2=0.0, k3=DPI, cte_rho0=1.0, cte_T0=1.0, cte_B02=0.0,
                      cte_B03=1.0, nodes=GAUSS_NODES, **overrides):
    s = Settings(gridpts=gridpts, geometry="Cartesian", k2=k2, k3=k3, **overrides)
    grid, xg = _grid(s.geometry, 0.0, 1.0, gridpts, nodes)
    one = np.ones_like(xg)
    fields = {"rho0": cte_rho0 * one, "T0": cte_T0 * one, "B02": cte_B02 * one,
              "B03": cte_B03 * one}
    return s, grid, xg, fields


def couette_flow_eq(gridpts=51, k2=0.0, k3=1.0, cte_rho0=1.0, cte_T0=1.0, cte_v02=0.0, cte_v03=1.0,
                    viscosity_value=1.0e-3, physics_type="hd", nodes=GAUSS_NODES, **overrides):
    """Plane Couette flow (flow + viscosity, Cartesian [0, 1]); the reference's HD regression case
    (tests/regression_tests/test_couette_flow_HD.py:24-50)."""
    s = Settings(gridpts=gridpts, geometry="Cartesian", k2=k2, k3=k3, physics_type=physics_type,
                 flow=True, viscosity=True, viscosity_value=viscosity_value, **overrides)
    grid, xg = _grid(s.geometry, 0.0, 1.0, gridpts, nodes)
    one = np.ones_like(xg)
    width = 1.0
    fields = {"rho0": cte_rho0 * one, "T0": cte_T0 * one,
              "v02": cte_v02 * xg / width, "dv02": cte_v02 / width * one,
              "v03": cte_v03 * xg / width, "dv03": cte_v03 / width * one}
    return s, grid, xg, fields


def suydam_cluster_eq(gridpts=51, k2=1.0, k3=-1.2, cte_rho0=1.0, cte_v02=0.0, cte_v03=0.14,
                      cte_p0=0.05, p1=0.1, alpha=2.0, nodes=GAUSS_NODES, **overrides):
    s = Settings(gridpts=gridpts, geometry="cylindrical", flow=True, k2=k2, k3=k3, **overrides)
    grid, r = _grid(s.geometry, 0.0, 1.0, gridpts, nodes)
    one = np.ones_like(r)
    J0, J1, J2 = jv(0, alpha * r), jv(1, alpha * r), jv(2, alpha * r)
    DJ0 = -alpha * J1
    DJ1 = alpha * (0.5 * J0 - 0.5 * J2)
    fields = {
        "rho0": cte_rho0 * one,
        "T0": (cte_p0 + 0.5 * p1 * J0**2) / cte_rho0,
        "dT0": p1 * J0 * DJ0 / cte_rho0,
        "v02": cte_v02 * one,
        "v03": cte_v03 * (1.0 - r**2),
        "dv03": -2.0 * cte_v03 * r,
        "B02": J1,
        "dB02": DJ1,
        "B03": np.sqrt(1.0 - p1) * J0,
        "dB03": -alpha * np.sqrt(1.0 - p1) * J1,
    }
    return s, grid, r, fields


def resistive_tearing_eq(gridpts=51, k2=0.49, k3=0.0, alpha=4.73884, beta=0.15, cte_rho0=1.0,
                         eta=1.0e-4, nodes=GAUSS_NODES, **overrides):
    s = Settings(gridpts=gridpts, geometry="Cartesian", resistivity=True, k2=k2, k3=k3,
                 **overrides)
    grid, x = _grid(s.geometry, -0.5, 0.5, gridpts, nodes)
    one = np.ones_like(x)
    B02 = np.sin(alpha * x)
    B03 = np.cos(alpha * x)
    B0 = np.sqrt(B02**2 + B03**2)
    fields = {
        "rho0": cte_rho0 * one,
        "T0": beta * B0 / 2.0,
        "B02": B02,
        "dB02": alpha * np.cos(alpha * x),
        "ddB02": -(alpha**2) * np.sin(alpha * x),
        "B03": B03,
        "dB03": -alpha * np.sin(alpha * x),
        "ddB03": -(alpha**2) * np.cos(alpha * x),
        "eta": eta * one,   # fixed resistivity: detadT = detadr = 0
    }
    return s, grid, x, fields


def magnetothermal_eq(gridpts=51, k2=0.0, k3=1.0, cte_T0=1.0, nodes=GAUSS_NODES, **overrides):
    s = Settings(gridpts=gridpts, geometry="cylindrical", cooling=True, heating=True,
                 conduction=True, perpendicular_conduction=False, k2=k2, k3=k3, **overrides)
    units = Units(unit_temperature=2.6e6, unit_magneticfield=10.0, unit_length=1.0e8,
                  mean_molecular_weight=1.0)
    grid, r = _grid(s.geometry, 0.0, 1.0, gridpts, nodes)
    one = np.ones_like(r)
    p0 = 1.0 / (2.0 * (1.0 + r**2) ** 2)
    rho0 = p0 / cte_T0
    drho0 = -2.0 * r / (cte_T0 * (r**2 + 1.0) ** 3)
    T0 = cte_T0 * one
    B02 = r / (1.0 + r**2)
    dB02 = (1.0 - r**2) / (r**4 + 2.0 * r**2 + 1.0)
    B0 = np.sqrt(B02**2)
    dB0 = (B02 * dB02) / B0
    lam = rosner_lambdaT(T0, units)
    dlam = rosner_dlambdadT(T0, units)
    kpara = tcpara(T0, units)
    dkpara = dtcparadT(T0, units)
    # thermal balance (mod_heatloss.f08:109-139): dT0 = ddT0 = v01 = B01 = tcperp = 0
    # here, so H = rho0 * lambdaT + (1 / rho0) * 0
    H = rho0 * lam + (1.0 / rho0) * (0.0 * one)
    # prefactor (mod_thermal_conduction.f08:217-247); dT0 = 0 -> dtcparadr = 0
    Kp = (kpara - 0.0) / B0**2
    dKp = ((0.0 - 0.0) * B0 - 2.0 * (kpara - 0.0) * dB0) / B0**3
    fields = {
        "rho0": rho0, "drho0": drho0, "T0": T0, "B02": B02, "dB02": dB02,
        "L0": rho0 * lam - H,
        "dLdT": rho0 * dlam,
        "dLdrho": lam,
        "tcpara": kpara, "dtcparadT": dkpara,
        "tcprefactor": Kp, "dtcprefactordr": dKp,
    }
    return s, grid, r, fields


def kelvin_helmholtz_cd_eq(gridpts=51, k2=-1.0, k3=None, V=1.63, cte_rho0=1.0, cte_p0=1.0,
                           Bz0=0.25, rc=0.5, rj=1.0, nodes=GAUSS_NODES, **overrides):
    Bth0 = 0.4 * (rc**2 + rj**2) / (rj * rc)
    a = 0.1 * rj
    if k3 is None:
        k3 = DPI / rj
    s = Settings(gridpts=gridpts, geometry="cylindrical", flow=True, k2=k2, k3=k3, **overrides)
    grid, r = _grid(s.geometry, 0.0, 2.0 * rj, gridpts, nodes)
    one = np.ones_like(r)
    rho0 = cte_rho0
    fields = {
        "rho0": rho0 * one,
        "T0": cte_p0 / rho0 - (Bth0**2 / (2.0 * rho0)) * (1.0 - rc**4 / (rc**2 + r**2) ** 2),
        "dT0": -(2.0 * Bth0**2 / rho0) * rc**4 * r / (r**2 + rc**2) ** 3,
        "v03": (V / 2.0) * np.tanh((rj - r) / a),
        "dv03": -(V / (2.0 * a)) / np.cosh((rj - r) / a) ** 2,
        "B02": Bth0 * r * rc / (rc**2 + r**2),
        "dB02": Bth0 * rc * (rc**2 - r**2) / (r**2 + rc**2) ** 2,
        "B03": Bz0 * one,
    }
    return s, grid, r, fields


def mri_accretion_eq(gridpts=5, k2=0.0, k3=70.0, beta=100.0, tau=1.0, nu=0.1,
                     x_start=1.0, x_end=2.0, nodes=GAUSS_NODES, **overrides):
    s = Settings(gridpts=gridpts, geometry="cylindrical", flow=True, gravity=True, k2=k2, k3=k3,
                 **overrides)
    grid, r = _grid(s.geometry, x_start, x_end, gridpts, nodes)
    mu1, epsilon = tau, nu
    p1 = epsilon**2
    Bz1 = np.sqrt(2.0 * p1 / (beta * (1.0 + mu1**2)))
    Bth1 = mu1 * Bz1
    vth1 = np.sqrt(1.0 - 2.5 * p1 - 0.25 * Bth1**2 - 1.25 * Bz1**2)
    rho0 = r ** (-1.5)
    drho0 = -1.5 * r ** (-2.5)
    p0 = p1 * r ** (-2.5)
    dp0 = -2.5 * p1 * r ** (-3.5)
    fields = {
        "rho0": rho0, "drho0": drho0,
        "T0": p0 / rho0,
        "dT0": (dp0 * rho0 - drho0 * p0) / rho0**2,
        "v02": vth1 / np.sqrt(r),
        "dv02": -0.5 * vth1 * r ** (-1.5),
        "B02": Bth1 * r ** (-1.25),
        "dB02": -1.25 * Bth1 * r ** (-2.25),
        "B03": Bz1 * r ** (-1.25),
        "dB03": -1.25 * Bz1 * r ** (-2.25),
        "g0": 1.0 / r**2,
    }
    return s, grid, r, fields


EQUILIBRIA = {
    "adiabatic_homo": adiabatic_homo_eq,
    "suydam_cluster": suydam_cluster_eq,
    "resistive_tearing": resistive_tearing_eq,
    "magnetothermal_instabilities": magnetothermal_eq,
    "kelvin_helmholtz_cd": kelvin_helmholtz_cd_eq,
    "MRI_accretion": mri_accretion_eq,
    "couette_flow": couette_flow_eq,
}
