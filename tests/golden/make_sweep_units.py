"""Workload definition of BASELINE config 5 (kelvin_helmholtz_cd, 256-point (k2, k3) wavenumber sweep): the list of
units (k2, k3, sigma) written to legolas_b200/data/sweep_config5_units.json.

k2 in {-3, -2, -1, 0} (integers: cylindrical geometry, src/mod_inspections.f08:117-129) x 64 values k3 = (j + 1) pi / 16,
the band of the plane where the current-driven Kelvin-Helmholtz mode is unstable.  The shift of a unit is what a Legolas
user would take: the most unstable eigenvalue of a coarse QR-invert run (G = 101, the reference's default solver,
src/solvers/smod_qr_invert.f08:46-135, restated in oracle.solvers.qr_invert), displaced by 2 % so that it is not an
eigenvalue of the fine pencil.  Each unit is then solved at G = 2001 with shift-invert Arnoldi, nev = 1.

Run here (CPU, ~3 min, single-threaded BLAS per worker):  OPENBLAS_NUM_THREADS=1 python tests/golden/make_sweep_units.py
"""
import json
import math
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

COARSE = 101
K2 = (-3, -2, -1, 0)
NK3 = 64


def coarse_mode(unit):
    from oracle import assembly as asm
    from oracle import equilibria as eq
    from oracle import solvers
    k2, k3 = unit
    s, grid, xg, f = eq.kelvin_helmholtz_cd_eq(gridpts=COARSE, k2=k2, k3=k3)
    A, B = asm.build_matrices(s, grid, xg, f)
    w = solvers.qr_invert(A.to_dense(), B.to_dense())
    w = w[np.isfinite(w) & (np.abs(w) < 1e6)]
    top = w[np.argmax(w.imag)]
    return [float(top.real), float(top.imag)]


def main():
    units = [(float(k2), math.pi * (j + 1) / 16.0) for k2 in K2 for j in range(NK3)]
    with ProcessPoolExecutor(max(1, (os.cpu_count() or 2) - 1)) as ex:
        modes = list(ex.map(coarse_mode, units, chunksize=2))
    out = []
    for (k2, k3), (re, im) in zip(units, modes):
        w0 = complex(re, im)
        sigma = w0 + 0.02 * abs(w0) * (0.6 + 0.8j)
        out.append({"k2": k2, "k3": k3, "coarse": [re, im], "sigma": [sigma.real, sigma.imag]})
    with open(os.path.join(os.path.dirname(os.path.dirname(HERE)), "legolas_b200", "data", "sweep_config5_units.json"), "w") as fh:
        json.dump({"equilibrium": "kelvin_helmholtz_cd", "coarse_gridpts": COARSE, "units": out}, fh, indent=0)
    print(len(out), "units")


if __name__ == "__main__":
    main()
