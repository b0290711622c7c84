"""The device path against the reference's OWN stored shift-invert runs for the optional-physics term groups
(SURVEY row A8): Hall, electron inertia, viscosity, viscous heating, resistivity, flow.  Needs a B200.

Each case is one of the reference's regression baselines (tests/regression_tests/baseline/BASE_*_SI_*.dat, setups
in test_uni_hall_adiabatic.py:87-94, test_uni_hall_elecinertia.py:71-95, test_taylor_couette.py:106-111,
test_uni_resistive.py:84-88, test_couette_flow.py:94-98, test_couette_flow_heating.py:95-99,
test_rotating_cylinder.py:44-48, test_rti_theta_pinch.py:39-43,89-93), extracted to tests/golden/*.npz by
make_golden.py.  The device assembles from the equilibrium arrays sampled on the baseline's (legacy, float32-rounded)
Gaussian grid, runs lgpu_shift_invert with the reference's defaults and must return the STORED eigenvalues to 1e-8
relative (north_star's tolerance) - no oracle in between; the oracle's matrices are compared as well.
"""
import numpy as np
import pytest

import legolas_b200 as lb
from oracle import assembly as asm
from test_oracle_golden import A8_PINS, LEGACY

pytestmark = pytest.mark.gpu


def to_host_settings(so: asm.Settings) -> lb.Settings:
    keys = ("gridpts", "geometry", "physics_type", "k2", "k3", "gamma", "incompressible", "flow",
            "resistivity", "cooling", "heating", "conduction", "perpendicular_conduction",
            "viscosity", "viscosity_value", "viscous_heating", "hall", "electron_inertia",
            "electron_fraction", "gravity", "boundary_type", "coaxial")
    s = lb.Settings(**{k: getattr(so, k) for k in keys})
    s.gauss_nodes, s.gauss_weights = so.gauss_nodes, so.gauss_weights
    return s


@pytest.mark.parametrize("name,eqf,sigma,nev,tol", A8_PINS)
def test_device_reproduces_reference_stored_eigenvalues(golden, name, eqf, sigma, nev, tol):
    g = golden(name)
    so, grid, xg, fields = eqf(gridpts=51, nodes=LEGACY[0])
    so.gauss_nodes, so.gauss_weights = LEGACY
    s = to_host_settings(so)
    s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=nev,
                                  which_eigenvalues="LM", sigma=sigma)
    ctx = lb.Context()
    try:
        mats = lb.build_matrices(s, grid, xg, fields, ctx=ctx)
        A, B = asm.build_matrices(so, grid, xg, fields)
        for M, label in ((A, "A"), (B, "B")):
            got = ctx.export_blocks(label)
            bound = 1e-12 * np.abs(M.blocks) + 1e-15 * np.abs(M.blocks).max()
            assert np.all(np.abs(got - M.blocks) <= bound), label
            r, c, v = ctx.export_coo(label)
            ro, co, vo = M.to_coo()
            assert np.array_equal(r, ro) and np.array_equal(c, co), label
        omega, vr, cfg, st = lb.solve_evp(mats, s)
        assert st["nconv"] == nev == len(g["eigenvalues"]), st
        worst = max(np.min(np.abs(omega - w)) / abs(w) for w in g["eigenvalues"])
        assert worst <= 1e-8, (name, worst)
        # eigenpairs of the pencil
        res = ctx.residuals(omega, np.asarray(vr))
        assert np.all(res <= 1e-8), res.max()
    finally:
        ctx.close()
