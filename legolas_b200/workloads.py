"""The two sharded workloads of BASELINE.json (configs 4 and 5) as lists of independent units, for
``bench.py``, ``scripts/sweep_bench.py`` and the GPU tests.  Host-side only: equilibrium sampling +
calls through the C ABI (``legolas_b200.api``); the reference runs such lists as one OS process per
parfile (post_processing/pylbo/automation/runner.py:202-211).

* ``SCAN_SHIFTS`` — config 4: magnetothermal_instabilities, G = 10 001, a 32-shift scan of the
  upper half plane next to the thermal branch.  Every shift was validated on
  the device under the reference's defaults (tol = 5e-15, maxiter = max(100, 10 nev),
  scripts/scan_probe.py): ``nev = 20`` where that converges, ``nev = 10`` where nev = 20 stalls at
  maxiter (SURVEY: "drop that shift to nev = 10"); shifts that stall with both (inside accumulation
  continua: Re 0 ... 0.006, Im 0.016 ... 0.028) are not part of the scan.  The third entry is the
  number of operator applications measured in that validation — the cost estimate used to hand the
  units out longest first (91 ... 251: the units are deliberately unequal).
* ``sweep_units()`` — config 5: kelvin_helmholtz_cd, G = 2001, 256 (k2, k3) points, per-unit shift
  from a coarse QR-invert pre-scan (tests/golden/make_sweep_units.py), nev = 1.
"""
from __future__ import annotations

import json
import math
import os
import warnings
from typing import Callable, List, Tuple

import numpy as np

from . import api, equilibria

HERE = os.path.dirname(os.path.abspath(__file__))

SCAN_GRIDPTS = 10001
# (sigma, nev, operator applications measured on the device)
SCAN_SHIFTS: List[Tuple[complex, int, int]] = [
    (0.000 + 0.010j, 20, 91), (0.006 + 0.010j, 20, 121), (0.012 + 0.010j, 20, 124), (0.018 + 0.010j, 20, 160),
    (0.024 + 0.010j, 20, 193), (0.012 + 0.016j, 20, 132), (0.018 + 0.016j, 20, 155), (0.024 + 0.016j, 20, 178),
    (0.012 + 0.022j, 20, 154), (0.018 + 0.022j, 20, 176), (0.024 + 0.022j, 20, 197), (0.018 + 0.028j, 20, 176),
    (0.024 + 0.028j, 20, 188), (0.018 + 0.034j, 20, 218), (0.024 + 0.034j, 20, 209), (0.018 + 0.040j, 20, 229),
    (0.024 + 0.040j, 20, 219), (0.024 + 0.046j, 20, 251), (-0.012 + 0.016j, 20, 132), (-0.024 + 0.016j, 20, 188),
    (-0.024 + 0.028j, 20, 199), (-0.024 + 0.040j, 20, 219),
    (0.012 + 0.028j, 10, 94), (0.006 + 0.034j, 10, 122), (0.012 + 0.034j, 10, 109), (0.000 + 0.040j, 10, 95),
    (0.006 + 0.040j, 10, 110), (0.012 + 0.040j, 10, 119), (0.006 + 0.046j, 10, 119), (0.012 + 0.046j, 10, 139),
    (-0.012 + 0.028j, 10, 95), (-0.012 + 0.040j, 10, 119),
]
SCAN_NEV_MAX = 20


def longest_first(costs) -> List[int]:
    """Unit indices by decreasing cost estimate (stable)."""
    return [int(i) for i in np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")]


class ScanSolver:
    """One library context holding the assembled config-4 matrices; ``solve(unit)`` factorises A - sigma B and
    runs the shift-invert Arnoldi iteration for one (sigma, nev, cost) unit.  Assembly happens once per context:
    the shifts of a scan share A and B."""

    def __init__(self, device: int = 0, gridpts: int = SCAN_GRIDPTS):
        self.settings, grid, fields = equilibria.magnetothermal_instabilities(gridpts)
        self.ctx = api.Context(device=device)
        self.settings.solvers = api.SolverSettings(solver="arnoldi", arpack_mode="shift-invert",
                                                   number_of_eigenvalues=SCAN_NEV_MAX, sigma=SCAN_SHIFTS[0][0])
        self.mats = api.build_matrices(self.settings, grid.base_grid, grid.gaussian_grid, fields, ctx=self.ctx)
        self.n_op = 0
        self.nconv_short = 0   # units that returned fewer pairs than asked for

    def __call__(self, unit) -> np.ndarray:
        sigma, nev, _ = unit
        self.settings.solvers = api.SolverSettings(solver="arnoldi", arpack_mode="shift-invert",
                                                   number_of_eigenvalues=nev, sigma=sigma)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            omega, _, _, st = api.solve_evp(self.mats, self.settings, want_vectors=False)
        self.n_op += st["n_op"]
        self.nconv_short += st["nconv"] < nev
        out = np.full(SCAN_NEV_MAX, np.nan + 1j * np.nan, dtype=np.complex128)
        out[:st["nconv"]] = omega[:st["nconv"]]
        return out

    def close(self):
        self.ctx.close()


# ------------------------------------------------------------------------------------ config 5
SWEEP_GRIDPTS = 2001
SWEEP_NEV, SWEEP_NCV, SWEEP_MAXITER = 1, 16, 20


def sweep_units(limit: int = 0) -> List[dict]:
    """The 256 units {k2, k3, sigma, coarse} of config 5; ``limit`` > 0 takes that many, evenly spread."""
    with open(os.path.join(HERE, "data", "sweep_config5_units.json")) as fh:
        units = json.load(fh)["units"]
    for i, u in enumerate(units):
        u["id"] = i
        u["sigma"] = complex(*u["sigma"])
        u["coarse"] = complex(*u["coarse"])
        u["cost"] = 0
    if 0 < limit < len(units):
        units = [units[i] for i in np.linspace(0, len(units) - 1, limit).astype(int)]
    return units


class SweepSolver:
    """One library context; ``solve(unit)`` samples the equilibrium of the unit's (k2, k3), assembles, factorises
    and tracks the mode next to the unit's shift (nev = 1, ncv = 16, at most 20 restarts)."""

    def __init__(self, device: int = 0, gridpts: int = SWEEP_GRIDPTS, sm_limit: int = 0):
        self.gridpts = gridpts
        self.ctx = api.Context(device=device)
        if sm_limit:
            self.ctx.set_sm_limit(sm_limit)   # share of the GPU when several units are in flight
        self.n_op = 0
        self.converged = 0

    def __call__(self, unit) -> np.ndarray:
        s, grid, fields = equilibria.kelvin_helmholtz_cd(self.gridpts, k2=unit["k2"], k3=unit["k3"])
        s.solvers = api.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=SWEEP_NEV,
                                       sigma=unit["sigma"], ncv=SWEEP_NCV, maxiter=SWEEP_MAXITER)
        mats = api.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=self.ctx)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)   # "maxiter reached": a unit without a mode next to its shift
            omega, _, _, st = api.solve_evp(mats, s, want_vectors=False)
        unit["cost"] = st["n_op"]                             # measured cost: orders a later pass longest first
        self.n_op += st["n_op"]
        self.converged += st["nconv"] >= SWEEP_NEV
        out = np.full(SWEEP_NEV, np.nan + 1j * np.nan, dtype=np.complex128)
        out[:st["nconv"]] = omega[:st["nconv"]]
        return out

    def close(self):
        self.ctx.close()


def make_factory(cls, device: int, **kw) -> Callable[[int], Callable]:
    return lambda worker: cls(device=device, **kw)


__all__ = ["SCAN_SHIFTS", "SCAN_NEV_MAX", "SCAN_GRIDPTS", "ScanSolver", "SweepSolver", "sweep_units", "longest_first",
           "make_factory", "SWEEP_GRIDPTS", "SWEEP_NEV", "SWEEP_NCV", "SWEEP_MAXITER", "math"]
