"""Scratch: where do shift-invert runs of the G=2001 Kelvin-Helmholtz sweep converge (reference defaults)?"""
import math, sys
import numpy as np
sys.path.insert(0, ".")
import legolas_b200 as lb
from legolas_b200 import equilibria as heq
ctx = lb.Context()
for k2, k3 in ((-1.0, math.pi), (-8.0, math.pi / 4), (3.0, 2 * math.pi), (7.0, 4 * math.pi), (0.0, math.pi / 2)):
    s, grid, fields = heq.kelvin_helmholtz_cd(2001, k2=k2, k3=k3)
    mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
    for s0 in (2.5 + 0.5j, 2.5 + 1.0j, 2.0 + 1.0j, 1.5 + 1.5j, 2.8 + 0.3j, 3.2 + 0.2j, 1.0 + 0.5j):
        sigma = s0 * k3 / math.pi
        s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=6, sigma=sigma)
        omega, _, _, st = lb.solve_evp(mats, s)
        near = omega[np.isfinite(omega)]
        print(f"k2={k2:5.1f} k3={k3:6.3f} s0={s0} nconv={st['nconv']} n_op={st['n_op']} "
              f"omega/k3*pi: {np.round(near[:3] * math.pi / k3, 3)}")
