"""Extract golden vectors from the reference's own datfiles into small ``.npz`` fixtures.

Run once in the build container (needs ``/root/reference``, which does not exist on
the GPU box):  ``python tests/golden/make_golden.py``

Sources (reference-owned data, never code):
  tests/regression_tests/baseline/BASE_*.dat      (Legolas 1.2.1 regression baselines)
  tests/pylbo_tests/utility_files/v2.0.0_mri_matrix.dat   (the only stored assembled A, B)
  tests/pylbo_tests/utility_files/v2.0.0_mri_subset_efs.dat (eigenvectors + eigenfunctions + residuals)
Each fixture keeps: eigenvalues, base grid, Gaussian grid, the stored equilibrium
arrays, parameters, units and (for the MRI file) the matrix triplets and the Gauss
nodes/weights printed in its header.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.datfile import read_datfile  # noqa: E402

REF = "/root/reference/tests"
FILES = {
    "uni_adiab_SI": "regression_tests/baseline/BASE_uni_adiab_SI_k2_0_k3_pi.dat",
    "uni_adiab_QR": "regression_tests/baseline/BASE_uni_adiab_QR_k2_0_k3_pi.dat",
    "suydam_QR": "regression_tests/baseline/BASE_suydam_QR_k2_1_k3_-1.2.dat",
    "resistive_tearing_QR": "regression_tests/baseline/BASE_resistive_tearing_QR_k2_0.49_k3_0.dat",
    "magnetothermal_SI": "regression_tests/baseline/BASE_magnetothermal_SI_k2_0_k3_1.dat",
    "magnetothermal_QR": "regression_tests/baseline/BASE_magnetothermal_QR_k2_0_k3_1.dat",
    "kh_cd_SI": "regression_tests/baseline/BASE_kelvin_helmholtz_current_driven_SI_k2_-1_k3_pi.dat",
    "kh_cd_QR": "regression_tests/baseline/BASE_kelvin_helmholtz_current_driven_QR_k2_-1_k3_pi.dat",
    # the reference's hydrodynamic ("hd", 5-variable state vector) regression case
    "couette_HD_QR": "regression_tests/baseline/BASE_couette_HD_QR_k2_0_k3_1.dat",
    # optional-physics term groups: Hall, electron inertia, viscosity, viscous heating, resistivity, flow
    "uni_adiab_hall_SI": "regression_tests/baseline/BASE_uni_adiab_hall_SI_k2_0.5pi_k3_0.87.dat",
    "uni_hall_elecinertia_SI": "regression_tests/baseline/BASE_uni_hall_elecinertia_SI_k2_5_k3_8.66.dat",
    "uni_hall_elecinertia_SI2": "regression_tests/baseline/BASE_uni_hall_elecinertia_SI2_k2_5_k3_8.66.dat",
    "taylor_couette_SI": "regression_tests/baseline/BASE_taylor_couette_SI_k2_0_k3_1.dat",
    "uni_resistive_SI": "regression_tests/baseline/BASE_uni_resistive_SI_k2_0_k3_1.dat",
    "couette_SI": "regression_tests/baseline/BASE_couette_SI_k2_0_k3_1.dat",
    "couette_heating_SI": "regression_tests/baseline/BASE_couette_heating_SI_k2_0_k3_1.dat",
    "rotating_cylinder_SI": "regression_tests/baseline/BASE_rotating_cylinder_SI_k2_1_k3_0.dat",
    "rti_theta_pinch_hd_SI": "regression_tests/baseline/BASE_rti_theta_pinch_hd_SI_k2_1_k3_0.dat",
    "rti_theta_pinch_mhd_SI": "regression_tests/baseline/BASE_rti_theta_pinch_mhd_SI_k2_1_k3_0.1.dat",
    "hall_harris_sheet_QR": "regression_tests/baseline/BASE_hall_harris_sheet_QR_k2_0.155_k3_0.01.dat",
    "mri_matrix": "pylbo_tests/utility_files/v2.0.0_mri_matrix.dat",
    # the only stored run with eigenvectors, eigenfunctions AND residuals (rows N1 of the scope table)
    "mri_subset_efs": "pylbo_tests/utility_files/v2.0.0_mri_subset_efs.dat",
}


def main():
    for key, rel in FILES.items():
        d = read_datfile(os.path.join(REF, rel))
        out = {
            "eigenvalues": d["eigenvalues"],
            "grid": d["grid"],
            "grid_gauss": d["grid_gauss"],
            "meta": json.dumps({
                "source": rel, "version": d["version"], "geometry": d["geometry"],
                "gridpoints": d["gridpoints"], "gamma": d["gamma"], "eq_type": d["eq_type"],
                "parameters": d["parameters"], "units": d["units"],
            }),
        }
        for name, arr in d["equilibria"].items():
            out["eq_" + name] = arr
        if "gauss_nodes" in d:
            out["gauss_nodes"] = d["gauss_nodes"]
            out["gauss_weights"] = d["gauss_weights"]
        if "eigenfunctions" in d:
            out["ef_grid"] = d["ef_grid"]
            out["ef_written_idxs"] = d["ef_written_idxs"]
            for name, arr in d["eigenfunctions"].items():
                out["ef_" + name] = arr
            out["state_vector"] = np.array(d["state_vector"])
        if "eigenvectors" in d:
            out["eigenvectors"] = d["eigenvectors"]
        if "residuals" in d:
            out["residuals"] = d["residuals"]
        if key == "mri_subset_efs":
            # the raw header of a 2.x datfile (everything before the eigenvalue block), the byte-level
            # pin of the writer in legolas_b200/datfile.py; plus the header entries the writer is fed
            from oracle import datfile as odf
            raw = open(os.path.join(REF, rel), "rb").read()
            st = odf._Stream(raw)
            st.s(len("legolas_version")), st.s(10), st.take("ii")
            odf._read_v2_header(st, {})
            out["header_bytes"] = np.frombuffer(raw[:st.pos], dtype=np.uint8)
            out["header_json"] = json.dumps({
                k: (v if not isinstance(v, complex) else [v.real, v.imag])
                for k, v in d.items()
                if k in ("x_start", "x_end", "ef_subset_radius", "ef_subset_center", "solver", "arpack_mode",
                         "number_of_eigenvalues", "which_eigenvalues", "ncv", "maxiter", "tolerance",
                         "boundary_type", "physics", "has_matrices", "has_eigenvectors", "has_residuals",
                         "has_efs", "has_derived_efs", "ef_subset_used")})
        if "matrix_A" in d:
            out["A_rows"], out["A_cols"], out["A_vals"] = d["matrix_A"]
            out["B_rows"], out["B_cols"], out["B_vals"] = d["matrix_B"]
        path = os.path.join(HERE, key + ".npz")
        np.savez_compressed(path, **out)
        print(f"{key}: {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
