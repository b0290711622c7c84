"""Row N2 of the scope table: the datfile writer (legolas_b200/datfile.py) against the reference's format.

CPU part: the header our writer produces for the reference's own stored run
(tests/pylbo_tests/utility_files/v2.0.0_mri_subset_efs.dat) is byte-identical to the stored one
except for the version string and the derived-eigenfunction flag, and a file with every optional
block switched on reads back block by block through the oracle's reader (which follows pylbo's).
The device-backed blocks (eigenfunctions, residuals, matrices) are served here by a stand-in built
on the oracle; tests/test_gpu_next_rows.py runs the same writer against the real context."""
import json

import numpy as np
import pytest

from legolas_b200 import datfile as ldf
from legolas_b200 import equilibria as heq
from legolas_b200.api import LegolasError, SolverSettings
from oracle import assembly as asm
from oracle import eigenfunctions as oef
from oracle import equilibria as oeq
from oracle.datfile import read_datfile


class OracleContext:
    """Stand-in for legolas_b200.Context in the CPU suite: same three calls, answered by the oracle."""

    def __init__(self, settings, base_grid, A, B, residuals):
        self.s, self.grid, self.A, self.B, self.res = settings, base_grid, A, B, residuals

    def eigenfunctions(self, vr, idxs):
        return oef.base_eigenfunctions(self.s.geometry, asm.STATE_VECTORS[self.s.physics_type], self.grid, vr,
                                       np.asarray(idxs) - 1)

    def residuals(self, omega, vr):
        return np.asarray(self.res, dtype=np.float64)

    def export_coo(self, which):
        return (self.A if which == "A" else self.B).to_coo()


def mri_run(golden):
    g = golden("mri_subset_efs")
    hdr = json.loads(str(g["header_json"]))
    meta = json.loads(str(g["meta"]))
    s, grid, fields = heq.mri_accretion(10)
    s.gauss_nodes, s.gauss_weights = g["gauss_nodes"], g["gauss_weights"]     # the stored run's (float32-rounded)
    s.solvers = SolverSettings(solver=hdr["solver"], arpack_mode=hdr["arpack_mode"],
                               number_of_eigenvalues=hdr["number_of_eigenvalues"],
                               which_eigenvalues=hdr["which_eigenvalues"], ncv=hdr["ncv"], maxiter=hdr["maxiter"],
                               tolerance=hdr["tolerance"])
    io = ldf.IoSettings(write_matrices=hdr["has_matrices"], write_eigenvectors=hdr["has_eigenvectors"],
                        write_residuals=hdr["has_residuals"], write_eigenfunctions=hdr["has_efs"],
                        write_ef_subset=hdr["ef_subset_used"], ef_subset_radius=hdr["ef_subset_radius"],
                        ef_subset_center=complex(*hdr["ef_subset_center"]))
    units = dict(meta["units"])
    info = ldf.RunInfo(grid_start=hdr["x_start"], grid_end=hdr["x_end"], equilibrium_type=meta["eq_type"],
                       cgs=bool(units.pop("cgs")), units=units, parameters=meta["parameters"])
    so, go, xgo, fo = oeq.mri_accretion_eq(gridpts=10)
    A, B = asm.build_matrices(so, go, xgo, fo)
    ctx = OracleContext(s, grid.base_grid, A, B, g["residuals"])
    return g, hdr, s, grid, fields, io, info, ctx


def test_header_bytes_match_reference_stored_run(tmp_path, golden):
    g, hdr, s, grid, fields, io, info, ctx = mri_run(golden)
    path = ldf.create_datfile(tmp_path / "mri.dat", s, grid.base_grid, grid.gaussian_grid, fields, g["eigenvalues"],
                              ctx=ctx, eigenvectors=g["eigenvectors"], io=io, info=info)
    ours = np.frombuffer(open(path, "rb").read(), dtype=np.uint8)
    ref = g["header_bytes"]
    ours = ours[:ref.size]
    diff = np.nonzero(ours != ref)[0]
    version = np.arange(15, 25)                       # "2.0.0" there, "2.0.6" here
    rest = np.setdiff1d(diff, version)
    # the only other difference: has_derived_efs (stored run: true; this library writes none)
    assert hdr["has_derived_efs"] is True
    assert rest.size == 1 and ref[rest[0]] == 1 and ours[rest[0]] == 0, rest
    assert bytes(ours[15:25]) == b"2.0.6     "


def test_selected_eigenfunction_subset_matches_reference(golden):
    g, hdr, s, grid, fields, io, info, ctx = mri_run(golden)
    flags, idxs = ldf.select_ef_subset(g["eigenvalues"], io)
    assert np.array_equal(idxs, g["ef_written_idxs"])
    assert flags.sum() == idxs.size
    assert np.allclose(ldf.ef_grid(grid.base_grid), g["ef_grid"], rtol=0, atol=1e-14)


def test_full_file_round_trip(tmp_path, golden):
    g, hdr, s, grid, fields, io, info, ctx = mri_run(golden)
    io.write_matrices = True
    path = ldf.create_datfile(tmp_path / "mri_all.dat", s, grid.base_grid, grid.gaussian_grid, fields,
                              g["eigenvalues"], ctx=ctx, eigenvectors=g["eigenvectors"], io=io, info=info)
    d = read_datfile(path)
    assert d["version"] == "2.0.6" and d["eq_type"] == "MRI_accretion" and d["geometry"] == "cylindrical"
    assert d["state_vector"] == list(asm.STATE_VECTORS["mhd"])
    assert d["dims"] == {"integralblock": 2, "subblock": 16, "quadblock": 32, "matrix": 160}
    assert d["gridpoints"] == 10 and d["gauss_gridpoints"] == 36 and d["ef_gridpoints"] == 19
    assert np.array_equal(d["gauss_nodes"], g["gauss_nodes"])
    assert d["physics"]["flow"] and d["physics"]["external_gravity"] and not d["physics"]["resistivity"]
    assert d["parameters"] == json.loads(str(g["meta"]))["parameters"]
    assert np.array_equal(d["eigenvalues"], g["eigenvalues"])
    assert np.array_equal(d["grid"], grid.base_grid) and np.array_equal(d["grid_gauss"], grid.gaussian_grid)
    # the 44 background arrays: this library's host sampling against what the reference stored
    assert list(d["equilibria"]) == list(ldf.BACKGROUND_NAMES)
    for name, arr in d["equilibria"].items():
        ref = g["eq_" + name]
        assert np.all(np.abs(arr - ref) <= 1e-12 * max(np.abs(ref).max(), 1e-300) + 1e-14), name
    assert np.array_equal(d["ef_written_idxs"], g["ef_written_idxs"])
    for name in d["state_vector"]:
        ref = g["ef_" + name]
        assert d["eigenfunctions"][name].shape == ref.shape
        assert np.all(np.abs(d["eigenfunctions"][name] - ref) <= 1e-12 * np.abs(ref).max()), name
    assert np.array_equal(d["eigenvectors"], g["eigenvectors"])
    assert np.array_equal(d["residuals"], g["residuals"])
    ra, ca, va = ctx.A.to_coo()
    rb, cb, vb = ctx.B.to_coo()
    assert np.array_equal(d["matrix_A"][0], ra) and np.array_equal(d["matrix_A"][1], ca)
    assert np.array_equal(d["matrix_A"][2], va)
    assert np.array_equal(d["matrix_B"][0], rb) and np.array_equal(d["matrix_B"][2], vb.real)


def test_minimal_file_and_argument_errors(tmp_path, golden):
    g, hdr, s, grid, fields, io, info, ctx = mri_run(golden)
    bare = ldf.IoSettings(write_background=False)
    path = ldf.create_datfile(tmp_path / "bare.dat", s, grid.base_grid, grid.gaussian_grid, fields,
                              g["eigenvalues"][:7], io=bare)
    d = read_datfile(path)
    assert d["equilibrium_names"] == [] and d["nb_eigenvalues"] == 7 and not d["has_efs"]
    assert np.array_equal(d["eigenvalues"], g["eigenvalues"][:7])
    with pytest.raises(LegolasError):      # eigenvectors requested, none given
        ldf.create_datfile(tmp_path / "x.dat", s, grid.base_grid, grid.gaussian_grid, fields, g["eigenvalues"],
                           io=ldf.IoSettings(write_eigenvectors=True))
    with pytest.raises(LegolasError):      # matrices requested without a context
        ldf.create_datfile(tmp_path / "x.dat", s, grid.base_grid, grid.gaussian_grid, fields, g["eigenvalues"],
                           io=ldf.IoSettings(write_matrices=True))
    with pytest.raises(LegolasError):      # wrong eigenvector shape
        ldf.create_datfile(tmp_path / "x.dat", s, grid.base_grid, grid.gaussian_grid, fields, g["eigenvalues"],
                           eigenvectors=np.zeros((3, 3), dtype=complex), io=ldf.IoSettings(write_eigenvectors=True))
    with pytest.raises(LegolasError):      # parameter the reference's file has no slot for
        ldf.create_datfile(tmp_path / "x.dat", s, grid.base_grid, grid.gaussian_grid, fields, g["eigenvalues"],
                           info=ldf.RunInfo(parameters={"not_a_parameter": 1.0}))
