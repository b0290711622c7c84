"""Minimal reader for Legolas datfiles (legacy <2.0 and 2.x layouts).

Test infrastructure (see ``oracle/__init__.py``).  pylbo cannot be imported in
this image (no matplotlib / f90nml), so the golden files of the reference are
parsed here.  Layout follows the reference's writer and pylbo's reader:

* writer: ``src/dataIO/mod_output.f08:56-120,476-508``
* v2 header: ``post_processing/pylbo/utilities/datfiles/header.py:53-338``
* legacy header: ``post_processing/pylbo/utilities/datfiles/header_legacy.py:24-210``
* data blocks: ``post_processing/pylbo/utilities/datfiles/file_reader.py:52-157``
"""
from __future__ import annotations

import struct

import numpy as np


class _Stream:
    def __init__(self, raw: bytes):
        self.raw = raw
        self.pos = 0

    def take(self, fmt: str):
        size = struct.calcsize("=" + fmt)
        out = struct.unpack_from("=" + fmt, self.raw, self.pos)
        self.pos += size
        return out

    def i(self) -> int:
        return self.take("i")[0]

    def b(self) -> bool:
        return bool(self.take("i")[0])

    def d(self) -> float:
        return self.take("d")[0]

    def z(self) -> complex:
        re, im = self.take("dd")
        return complex(re, im)

    def s(self, length: int) -> str:
        out = self.raw[self.pos:self.pos + length].decode("ascii", "replace")
        self.pos += length
        return out.strip()

    def strs(self, length: int, amount: int) -> list:
        return [self.s(length) for _ in range(amount)]

    def arr(self, dtype, amount: int) -> np.ndarray:
        dt = np.dtype(dtype)
        out = np.frombuffer(self.raw, dtype=dt, count=amount, offset=self.pos).copy()
        self.pos += dt.itemsize * amount
        return out


def _version_tuple(version: str):
    return tuple(int(p) for p in version.split(".")[:3])


def read_datfile(path) -> dict:
    """Parse a datfile into a plain dict.

    Keys: version, geometry, gridpoints, gauss_gridpoints, gamma, eq_type,
    parameters, units, equilibrium_names, eigenvalues, grid, grid_gauss,
    equilibria (dict name -> array) and, when present, gauss_nodes /
    gauss_weights (v2 only), matrix_A / matrix_B triplets (1-based rows, cols).
    Eigenfunction blocks are skipped.
    """
    with open(path, "rb") as fh:
        st = _Stream(fh.read())
    tag = st.s(len("legolas_version"))
    if tag != "legolas_version":
        raise ValueError(f"{path}: unsupported (pre-1.0) datfile")
    version = st.s(10)
    vt = _version_tuple(version)
    data = {"version": version}
    str_len, str_len_arr = st.take("ii")
    if vt >= (2, 0, 0):
        _read_v2_header(st, data)
    else:
        _read_legacy_header(st, data, vt, str_len, str_len_arr)
    _read_blocks(st, data, vt, str_len_arr)
    return data


def _read_legacy_header(st, data, vt, str_len, str_len_arr):
    data["geometry"] = st.s(str_len)
    data["x_start"], data["x_end"] = st.take("dd")
    (data["gridpoints"], data["gauss_gridpoints"], data["matrix_gridpoints"],
     data["ef_gridpoints"]) = st.take("iiii")
    data["gamma"] = st.d()
    data["eq_type"] = st.s(str_len)
    data["has_efs"] = st.b()
    data["has_derived_efs"] = st.b() if vt >= (1, 1, 3) else False
    data["has_matrices"] = st.b()
    data["has_eigenvectors"] = st.b() if vt >= (1, 3, 0) else False
    data["has_residuals"] = st.b() if vt >= (1, 3, 0) else False
    if vt >= (1, 1, 4):
        data["ef_subset_used"] = st.b()
        data["ef_subset_center"] = st.z()
        data["ef_subset_radius"] = st.d()
    nb_params = st.i()
    len_name = st.i() if vt >= (1, 0, 2) else str_len_arr
    names = st.strs(len_name, nb_params)
    values = st.arr("f8", nb_params)
    data["parameters"] = {n: float(v) for n, v in zip(names, values) if not np.isnan(v)}
    nb_names = st.i()
    len_name = st.i() if vt >= (1, 0, 2) else str_len_arr
    data["equilibrium_names"] = st.strs(len_name, nb_names)
    units = {"cgs": st.b()}
    nb_units, len_unit = st.take("ii")
    unit_names = st.strs(len_unit, nb_units)
    unit_values = st.arr("f8", nb_units)
    units.update({n: float(v) for n, v in zip(unit_names, unit_values)})
    data["units"] = units
    data["nb_eigenvalues"] = st.i()
    data["nb_eqs"] = 8
    data["state_vector"] = ["rho", "v1", "v2", "v3", "T", "a1", "a2", "a3"]


def _read_v2_header(st, data):
    data["nb_eqs"] = st.i()
    data["physics_type"] = st.s(st.i())
    len_name, size_vector = st.take("ii")
    data["state_vector"] = st.strs(len_name, size_vector)
    data["dims"] = dict(zip(("integralblock", "subblock", "quadblock", "matrix"), st.take("iiii")))
    data["geometry"] = st.s(st.i())
    data["gridpoints"], data["gauss_gridpoints"], data["ef_gridpoints"] = st.take("iii")
    n_gauss = st.i()
    data["gauss_nodes"] = st.arr("f8", n_gauss)
    data["gauss_weights"] = st.arr("f8", n_gauss)
    data["x_start"], data["x_end"] = st.take("dd")
    for key in ("has_matrices", "has_eigenvectors", "has_residuals", "has_efs",
                "has_derived_efs", "ef_subset_used"):
        data[key] = st.b()
    data["ef_subset_radius"] = st.d()
    data["ef_subset_center"] = st.z()
    data["solver"] = st.s(st.i())
    data["arpack_mode"] = st.s(st.i())
    data["number_of_eigenvalues"] = st.i()
    data["which_eigenvalues"] = st.s(st.i())
    data["ncv"] = st.i()
    data["maxiter"] = st.i()
    data["sigma"] = st.z()
    data["tolerance"] = st.d()
    data["eq_type"] = st.s(st.i())
    data["boundary_type"] = st.s(st.i())
    n_units = st.i()
    units = {"cgs": st.b()}
    for _ in range(n_units):
        name = st.s(st.i())
        units[name] = st.d()
    data["units"] = units
    data["gamma"] = st.d()
    data["is_incompressible"] = st.b()
    physics = {}
    physics["flow"] = st.b()
    physics["cooling"] = st.b()
    physics["cooling_curve"] = st.s(st.i())
    physics["interpolation_points"] = st.i()
    for key in ("external_gravity", "resistivity", "has_fixed_resistivity", "viscosity",
                "has_viscous_heating", "conduction", "has_parallel_conduction",
                "has_fixed_tc_para", "has_perpendicular_conduction", "has_fixed_tc_perp",
                "Hall", "Hall_uses_substitution", "has_electron_inertia"):
        physics[key] = st.b()
    data["physics"] = physics
    nb_params, len_name = st.take("ii")
    params = {}
    for _ in range(nb_params):
        name = st.s(len_name)
        params[name] = st.d()
    data["parameters"] = {k: v for k, v in params.items() if not np.isnan(v)}
    nb_names, len_name = st.take("ii")
    names = st.strs(len_name, nb_names) if nb_names > 0 else []
    data["equilibrium_names"] = [n.replace("db03", "dB03") for n in names]
    data["nb_eigenvalues"] = st.i()


def _read_blocks(st, data, vt, str_len_arr):
    nev = data["nb_eigenvalues"]
    data["eigenvalues"] = st.arr("c16", nev)
    data["grid"] = st.arr("f8", data["gridpoints"])
    data["grid_gauss"] = st.arr("f8", data["gauss_gridpoints"])
    data["equilibria"] = {
        name: st.arr("f8", data["gauss_gridpoints"]) for name in data["equilibrium_names"]
    }
    nb_written = nev
    if data.get("has_efs"):
        if vt >= (2, 0, 0):
            ef_gridsize = st.i()
            nb_ef_names = data["nb_eqs"]
        else:
            nb_ef_names = st.i()
            st.strs(str_len_arr, nb_ef_names)
            ef_gridsize = data["ef_gridpoints"]
        data["ef_grid"] = st.arr("f8", ef_gridsize)
        if vt >= (1, 1, 4):
            flags = st.arr("i4", st.i())
            idxs = st.arr("i4", st.i())
            nb_written = len(idxs)
            data["ef_written_idxs"] = idxs          # 1-based eigenvalue indices (mod_output.f08:410-411)
            del flags
        # one (ef_gridpts, nb_written) column-major block per state-vector entry (mod_output.f08:412-414)
        if vt >= (2, 0, 0):
            npts = data["ef_gridpoints"]
            data["eigenfunctions"] = {
                name: st.arr("c16", npts * nb_written).reshape((npts, nb_written), order="F")
                for name in data["state_vector"]
            }
        else:
            st.pos += 16 * data["ef_gridpoints"] * nb_written * nb_ef_names
    if data.get("has_derived_efs"):
        if vt >= (2, 0, 0):
            nb_names, size_names = st.take("ii")
        else:
            nb_names, size_names = st.i(), str_len_arr
        st.strs(size_names, nb_names)
        st.pos += 16 * data["ef_gridpoints"] * nb_written * nb_names
    if data.get("has_eigenvectors"):
        length, count = st.take("ii")
        data["eigenvectors"] = st.arr("c16", length * count).reshape((length, count), order="F")
    if data.get("has_residuals"):
        data["residuals"] = st.arr("f8", st.i())
    if data.get("has_matrices"):
        nnz_b, nnz_a = st.take("ii")
        rec_b = st.arr(np.dtype([("r", "<i4"), ("c", "<i4"), ("v", "<f8")]), nnz_b)
        rec_a = st.arr(np.dtype([("r", "<i4"), ("c", "<i4"), ("re", "<f8"), ("im", "<f8")]), nnz_a)
        data["matrix_B"] = (rec_b["r"].astype(np.int64), rec_b["c"].astype(np.int64),
                            rec_b["v"].astype(np.float64))
        data["matrix_A"] = (rec_a["r"].astype(np.int64), rec_a["c"].astype(np.int64),
                            rec_a["re"] + 1j * rec_a["im"])
