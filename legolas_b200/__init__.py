"""legolas_b200 — B200-native hot path of Legolas: finite-element assembly of A and B plus
the shift-invert Arnoldi eigen-solve, behind a C ABI (include/legolas_b200.h).

``legolas_b200.api`` mirrors the reference's host interface (build_matrices / solve_evp),
``legolas_b200.equilibria`` samples the benchmark equilibria on the host,
``legolas_b200.sweep`` shards independent shifts / wavenumbers over the GPUs of one node,
``legolas_b200.datfile`` writes the reference's datfile from device-resident results.
"""
from .api import (ArpackConfig, Context, LegolasError, Matrices, Settings, SolverSettings,  # noqa: F401
                  build_matrices, new_arpack_config, solve_evp, zlarnv)
from ._lib import LgpuError  # noqa: F401

__version__ = "0.1.0"
