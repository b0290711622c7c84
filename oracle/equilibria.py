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
  resistive_homo ............ src/equilibria/smod_equil_resistive_homo.f08:26-67
  taylor_couette ............ src/equilibria/smod_equil_taylor_couette.f08:31-107
  rotating_plasma_cylinder .. src/equilibria/smod_equil_rotating_plasma_cylinder.f08:29-120
  RTI_theta_pinch ........... src/equilibria/smod_equil_RTI_theta_pinch.f08:33-119
  harris_sheet .............. src/equilibria/smod_equil_harris_sheet.f08:26-128
  Hall / inertia factors .... src/physics/mod_hall.f08:45-84
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
ME_CGS = 9.1094e-28          # src/mod_physical_constants.f08:25
EC_CGS = 4.8032e-10          # src/mod_physical_constants.f08:29
TC_PF_KAPPA_PARA = 1.8e-5
TC_PF_KAPPA_PERP = 8.2e-13

LOGT_ROSNER = np.array([3.89063, 4.30195, 4.575, 4.9, 5.4, 5.77, 6.315, 7.60457])
LOGXI_ROSNER = np.array([-69.900, -48.307, -21.850, -31.000, -21.200, -10.400,
                         -21.940, -17.730, -26.602])
ALPHA_ROSNER = np.array([11.7, 6.15, 0.0, 2.0, 0.0, -2.0, 0.0, -0.666666667, 0.5])


class Units:
    """units_t set from temperature (mod_units.f08:118-134,161-211)."""

    def __init__(self, unit_length=1.0e9, unit_magneticfield=10.0, unit_temperature=1.0e6,
                 mean_molecular_weight=0.5, unit_density=None):
        self.unit_length = unit_length
        self.unit_magneticfield = unit_magneticfield
        self.mean_molecular_weight = mean_molecular_weight
        self.unit_pressure = unit_magneticfield**2 / MU0_CGS
        if unit_density is not None:   # set_units_from_density (mod_units.f08:96-115,176-186)
            self.unit_density = unit_density
            unit_temperature = mean_molecular_weight * self.unit_pressure * MP_CGS / (KB_CGS * unit_density)
        else:
            self.unit_density = (
                mean_molecular_weight * self.unit_pressure * MP_CGS / (KB_CGS * unit_temperature)
            )
        self.unit_temperature = unit_temperature
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
def hall_factors(units: Units, inertia: bool):
    """mod_hall.f08:45-84 without drop-off profiles: (hallfactor, inertiafactor)."""
    hf = (MP_CGS * units.unit_velocity) / (EC_CGS * units.unit_length * units.unit_magneticfield)
    inf = (MP_CGS * ME_CGS * units.unit_velocity**2
           / (EC_CGS * units.unit_length * units.unit_magneticfield) ** 2) if inertia else 0.0
    return hf, inf


def adiabatic_homo_eq(gridpts=51, k2=0.0, k3=DPI, cte_rho0=1.0, cte_T0=1.0, cte_B02=0.0,
                      cte_B03=1.0, x_start=0.0, x_end=1.0, units: Units | None = None, nodes=GAUSS_NODES,
                      **overrides):
    """With ``hall=True`` (and ``electron_inertia``) in the overrides this is the reference's uniform Hall case
    (tests/regression_tests/test_uni_hall_adiabatic.py:9-31, test_uni_hall_elecinertia.py:9-29)."""
    s = Settings(gridpts=gridpts, geometry="Cartesian", k2=k2, k3=k3, **overrides)
    grid, xg = _grid(s.geometry, x_start, x_end, gridpts, nodes)
    one = np.ones_like(xg)
    fields = {"rho0": cte_rho0 * one, "T0": cte_T0 * one, "B02": cte_B02 * one,
              "B03": cte_B03 * one}
    if s.hall:
        hf, inf = hall_factors(units or Units(), s.electron_inertia)
        fields["hallfactor"] = hf * one
        if s.electron_inertia:
            fields["inertiafactor"] = inf * one
    return s, grid, xg, fields


HALL_UNITS = dict(unit_length=7.534209349981049e-9, unit_magneticfield=10.0, unit_density=1.7e-14,
                  mean_molecular_weight=1.0)   # test_uni_hall_adiabatic.py:25-30


def uni_hall_eq(gridpts=51, inertia=False, k=DPI, nodes=GAUSS_NODES, **overrides):
    """Uniform medium with Hall (and electron inertia) terms on [0, 1000]: k = pi rounded to 14 digits per component
    for the adiabatic case (test_uni_hall_adiabatic.py:14-15), k = 10 for the inertia case
    (test_uni_hall_elecinertia.py:14-15), both at 30 degrees to B."""
    if inertia:
        k2, k3 = 10 * np.sin(np.pi / 6), 10 * np.cos(np.pi / 6)
    else:
        k2, k3 = round(np.pi * np.sin(np.pi / 6), 14), round(np.pi * np.cos(np.pi / 6), 14)
    return adiabatic_homo_eq(gridpts=gridpts, k2=k2, k3=k3, x_start=0.0, x_end=1000.0, units=Units(**HALL_UNITS),
                             nodes=nodes, hall=True, electron_inertia=inertia, electron_fraction=0.5, **overrides)


def resistive_homo_eq(gridpts=51, k2=0.0, k3=1.0, beta=0.25, cte_rho0=1.0, cte_B02=0.0, cte_B03=1.0,
                      eta=1.0e-3, nodes=GAUSS_NODES, **overrides):
    s = Settings(gridpts=gridpts, geometry="Cartesian", resistivity=True, k2=k2, k3=k3, **overrides)
    grid, xg = _grid(s.geometry, 0.0, 1.0, gridpts, nodes)
    one = np.ones_like(xg)
    B0 = np.sqrt(cte_B02**2 + cte_B03**2)
    fields = {"rho0": cte_rho0 * one, "T0": beta * B0**2 / 2.0 * one, "B02": cte_B02 * one,
              "B03": cte_B03 * one, "eta": eta * one}
    return s, grid, xg, fields


def taylor_couette_eq(gridpts=51, k2=0.0, k3=1.0, cte_rho0=1.0, alpha=1.0, beta=2.0, x_start=1.0, x_end=2.0,
                      viscosity_value=1.0e-3, nodes=GAUSS_NODES, **overrides):
    s = Settings(gridpts=gridpts, geometry="cylindrical", flow=True, coaxial=True, viscosity=True,
                 viscosity_value=viscosity_value, k2=k2, k3=k3, **overrides)
    grid, r = _grid(s.geometry, x_start, x_end, gridpts, nodes)
    Rrat = x_start / x_end
    A = (alpha * Rrat**2 - beta) / (Rrat**2 - 1.0)
    B = x_start**2 * (alpha - beta) / (1.0 - Rrat**2)
    Tstart = 0.5 * ((A * x_start) ** 2 + 4.0 * A * B * np.log(x_start) - (B / x_start) ** 2)
    T0 = 0.5 * ((A * r) ** 2 + 4.0 * A * B * np.log(r) - (B / r) ** 2)
    if not Tstart > 0.0:
        T0 = 2.0 * abs(Tstart) + T0
    v02 = A * r + B / r
    fields = {"rho0": cte_rho0 * np.ones_like(r), "T0": T0, "dT0": v02**2 / r,
              "v02": v02, "dv02": A - B / r**2, "ddv02": 2.0 * B / r**3}
    return s, grid, r, fields


def rotating_plasma_cylinder_eq(gridpts=51, k2=1.0, k3=0.0, p1=8.0, p2=0.0, p3=0.0, p4=1.0, p5=0.0, p6=0.0,
                                cte_p0=0.1, cte_rho0=1.0, nodes=GAUSS_NODES, **overrides):
    s = Settings(gridpts=gridpts, geometry="cylindrical", flow=True, k2=k2, k3=k3, **overrides)
    grid, r = _grid(s.geometry, 0.0, 1.0, gridpts, nodes)
    one = np.ones_like(r)
    a21, a22, a3, b21, b22, b3 = p1, p2, p3, p4, p5, p6
    rho0 = cte_rho0
    fields = {
        "rho0": rho0 * one,
        "T0": (1.0 / rho0) * (cte_p0 + 0.5 * (a21**2 - 2.0 * b21**2) * r**2
                              + (2.0 / 3.0) * (a21 * a22 - b21 * b22) * r**3
                              + (1.0 / 4.0) * (a22**2 - b22**2) * r**4),
        "dT0": (1.0 / rho0) * ((a21**2 - 2.0 * b21**2) * r + 2.0 * (a21 * a22 - b21 * b22) * r**2
                               + (a22**2 - b22**2) * r**3),
        "v02": a21 * r + a22 * r**2, "dv02": a21 + 2.0 * a22 * r, "v03": a3 * one,
        "B02": b21 * r + b22 * r**2, "dB02": b21 + 2.0 * b22 * r, "B03": b3 * one,
    }
    return s, grid, r, fields


def rti_theta_pinch_eq(gridpts=51, k2=1.0, k3=0.0, cte_rho0=1.0, alpha=2.0, delta=1.0 / 6.0, r0=0.0,
                       nodes=GAUSS_NODES, **overrides):
    s = Settings(gridpts=gridpts, geometry="cylindrical", flow=True, k2=k2, k3=k3, **overrides)
    grid, r = _grid(s.geometry, 0.0, 1.0, gridpts, nodes)
    width = grid[-1] - grid[0]          # after the on-axis shift of the grid start
    cte_p0 = 0.5 * (1.0 - delta) ** 2
    B_inf = width * np.sqrt(cte_rho0)
    bigO = alpha * np.sqrt(2.0 * delta * (1.0 - delta))
    x = r / width
    fx = alpha**2 * (x**2 - r0**2)
    dfx = alpha**2 * 2.0 * x / width
    fields = {
        "rho0": cte_rho0 / np.cosh(fx) ** 2,
        "drho0": -2.0 * cte_rho0 * dfx * np.tanh(fx) / np.cosh(fx) ** 2,
        "T0": cte_p0 / cte_rho0 * np.ones_like(r),
        "v02": bigO * r, "dv02": bigO * np.ones_like(r),
        "B03": B_inf * (delta + (1.0 - delta) * np.tanh(fx)),
        "dB03": B_inf * (1.0 - delta) * dfx / np.cosh(fx) ** 2,
    }
    return s, grid, r, fields


def harris_sheet_eq(gridpts=51, k2=0.155, k3=0.01, alpha=1.0, cte_rho0=1.0, cte_B02=1.0, cte_B03=5.0,
                    eta=1.0e-4, nodes=GAUSS_NODES, **overrides):
    """The reference's Hall regression case (tests/regression_tests/test_hall_harris_sheet.py:8-33): resistive,
    Hall, incompressible, eq_bool = .false. branch of the equilibrium."""
    kw = dict(resistivity=True, hall=True, electron_fraction=0.5, incompressible=True)
    kw.update(overrides)
    s = Settings(gridpts=gridpts, geometry="Cartesian", k2=k2, k3=k3, **kw)
    grid, x = _grid(s.geometry, -15.0, 15.0, gridpts, nodes)
    one = np.ones_like(x)
    B02 = cte_B02 * np.tanh(x / alpha)
    B03 = cte_B03 * one
    B0 = np.sqrt(B02**2 + B03**2)
    fields = {
        "rho0": cte_rho0 * one,
        "T0": (cte_B03**2 + cte_B02**2 - B0**2) / (2.0 * cte_rho0),
        "dT0": -cte_B02**2 * np.sinh(x / alpha) / (alpha * cte_rho0 * np.cosh(x / alpha) ** 3),
        "B02": B02, "dB02": cte_B02 / (alpha * np.cosh(x / alpha) ** 2),
        "ddB02": -2.0 * cte_B02 * np.sinh(x / alpha) / (alpha**2 * np.cosh(x / alpha) ** 3),
        "B03": B03, "eta": eta * one,
    }
    if s.hall:
        fields["hallfactor"] = hall_factors(Units(**HALL_UNITS), False)[0] * one
    return s, grid, x, fields


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
    "resistive_homo": resistive_homo_eq,
    "taylor_couette": taylor_couette_eq,
    "rotating_plasma_cylinder": rotating_plasma_cylinder_eq,
    "RTI_theta_pinch": rti_theta_pinch_eq,
    "harris_sheet": harris_sheet_eq,
    "uni_hall": uni_hall_eq,
}
