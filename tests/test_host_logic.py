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


def test_cgs2_ring_protocol_model():
    """Model of the shared-memory ring of the fused Gram-Schmidt step (legolas_b200/csrc/arnoldi.cu,
    cgs_fill_number / load_tile / the producer lane): local tile i lives in slot i % S, the three
    passes visit the CTA's tiles up / down / up, the S tiles a pass ends with stay resident for the
    next one.  Checked here for every shape: per slot the producer's fill numbers are 0, 1, 2, ... in
    issue order (they select the mbarrier phase), a fill is only issued after the release that
    precedes it, every visit finds its own tile in its slot, and nothing deadlocks.  The device code
    evaluates the same formulas with a multiply-shift division, checked as well."""
    def fill_number(p, i, nt, S):
        sg = i % S
        n = (nt - sg + S - 1) // S
        if p == 1:
            return i // S
        if p == 2:
            return n + (sg + (n - 1) * S - S - i) // S
        return 2 * n - 1 + i // S - 1

    for S in range(2, 9):
        M = 65536 // S + 1
        assert all((x * M) >> 16 == x // S for x in range(8192))
        for nt in range(1, 48):
            fills = ([(1, i) for i in range(nt)] + [(2, i) for i in range(nt - S - 1, -1, -1)]
                     + [(3, i) for i in range(S, nt)])
            issued, number = {}, {}
            for p, i in fills:
                F = fill_number(p, i, nt, S)
                assert issued.get(i % S, 0) == F, (S, nt, p, i)
                issued[i % S] = F + 1
                number[(p, i)] = F
            pending, content, releases = list(fills), {}, {}

            def produce():
                while pending:
                    p, i = pending[0]
                    F = number[(p, i)]
                    if F >= 1 and releases.get(i % S, 0) < F:
                        return
                    content[i % S] = (i, F)
                    pending.pop(0)

            produce()
            visits = ([(1, i) for i in range(nt)] + [(2, i) for i in range(nt - 1, -1, -1)]
                      + [(3, i) for i in range(nt)])
            for p, i in visits:
                sg = i % S
                resident = (p == 2 and i >= nt - S) or (p == 3 and i < S)
                if not resident:
                    produce()
                    assert content.get(sg, (None, -1))[1] == fill_number(p, i, nt, S), (S, nt, p, i)
                assert content[sg][0] == i, (S, nt, p, i)
                keep = (p == 1 and i >= nt - S) or (p == 2 and i < S)
                if not keep:
                    releases[sg] = releases.get(sg, 0) + 1
                    produce()
            assert not pending
