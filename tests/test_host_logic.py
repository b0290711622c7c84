"""CPU tests of the host side: C-ABI surface, zlarnv port, arpack_t defaults/validation
(tests/unit_tests/mod_test_solvers_arpack_type.pf), host equilibrium sampling vs the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

import legolas_b200 as lb
from legolas_b200 import _lib, equilibria as heq
from oracle import equilibria as oeq
from oracle import solvers as osolvers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "legolas_b200.h")).read()
    declared = set(re.findall(r"\b(lgpu_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    _lib.load()


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.CSettings) == 18 * 4 + 5 * 8 + 8 * 8
    assert ctypes.sizeof(_lib.CArnoldi) == 3 * 4 + 4 + 3 * 8 + 2 * 4
    assert ctypes.sizeof(_lib.CStats) == 8 * 4 + 3 * 8


def test_zlarnv_port_is_bit_exact():
    for n in (1, 63, 64, 65, 816, 5000):
        assert np.array_equal(lb.zlarnv(n), osolvers.zlarnv(n))


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lb.LgpuError) as err:
        lb.Context()
    assert err.value.code == _lib.ENOGPU


# ---- mod_test_solvers_arpack_type.pf
def test_arpack_config_defaults():
    sv = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=10)
    cfg = lb.new_arpack_config(100, 2, "I", sv)
    assert (cfg.ncv, cfg.maxiter, cfg.tolerance) == (20, 100, 5.0e-15)
    assert (sv.ncv, sv.maxiter) == (20, 100)          # written back (mod_arpack_type.f08:98-99)
    assert cfg.info == 1 and cfg.iparam[1] == 1 and cfg.iparam[3] == 100 and cfg.iparam[7] == 2
    sv = lb.SolverSettings(number_of_eigenvalues=20)
    assert lb.new_arpack_config(1000, 2, "I", sv).maxiter == 200
    sv = lb.SolverSettings(number_of_eigenvalues=8)
    assert lb.new_arpack_config(10, 2, "I", sv).ncv == 10   # min(2 nev, N)


@pytest.mark.parametrize("kw,msg", [
    (dict(number_of_eigenvalues=0), "number of eigenvalues"),
    (dict(number_of_eigenvalues=100), "matrix size"),
    (dict(number_of_eigenvalues=5, ncv=5), "ncv too low"),
    (dict(number_of_eigenvalues=5, ncv=101), "ncv too high"),
    (dict(number_of_eigenvalues=5, maxiter=-1), "maxiter"),
    (dict(number_of_eigenvalues=5, which_eigenvalues="XX"), "which_eigenvalues"),
])
def test_arpack_config_validation(kw, msg):
    with pytest.raises(lb.LegolasError, match=msg):
        lb.new_arpack_config(100, 2, "I", lb.SolverSettings(**kw))
    with pytest.raises(lb.LegolasError):
        lb.new_arpack_config(100, 4, "I", lb.SolverSettings())
    with pytest.raises(lb.LegolasError):
        lb.new_arpack_config(100, 2, "X", lb.SolverSettings())


PAIRS = [
    ("adiabatic_homo", {}), ("suydam_cluster", {}), ("resistive_tearing", {}),
    ("magnetothermal_instabilities", {}), ("kelvin_helmholtz_cd", {}), ("MRI_accretion", {}),
    ("magnetothermal_instabilities", {"k2": 10.0}),
]


@pytest.mark.parametrize("name,kw", PAIRS)
def test_host_sampling_matches_oracle(name, kw):
    s, grid, fields = heq.EQUILIBRIA[name](gridpts=31, **kw)
    so, go, xgo, fo = oeq.EQUILIBRIA[name](gridpts=31, **kw)
    assert np.array_equal(grid.base_grid, go)
    assert np.abs(grid.gaussian_grid - xgo).max() < 1e-15
    for key in ("geometry", "k2", "k3", "flow", "resistivity", "cooling", "heating", "conduction",
                "perpendicular_conduction", "gravity"):
        assert getattr(s, key) == getattr(so, key), key
    assert set(fields) == set(fo)
    for key, ref in fo.items():
        assert np.abs(fields[key] - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max()), key
