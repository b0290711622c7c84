"""ctypes binding of liblegolas_b200.so (include/legolas_b200.h).

There is no CPU fallback: importing works anywhere (so the host logic can be tested), but
every compute entry point needs the compiled library and a CUDA device and fails loudly
otherwise.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LGPU_LIB") or os.path.join(HERE, "liblegolas_b200.so")

N_FIELDS = 38

OK, EINVAL, ENOGPU, ESTATE, ENOMEM = 0, -1, -2, -3, -4


class LgpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"legolas_b200 error {code}: {message}")
        self.code = code


class CSettings(C.Structure):
    _fields_ = [
        ("gridpts", C.c_int32), ("physics_type", C.c_int32), ("geometry", C.c_int32),
        ("incompressible", C.c_int32), ("flow", C.c_int32), ("resistivity", C.c_int32),
        ("cooling", C.c_int32), ("heating", C.c_int32), ("conduction", C.c_int32),
        ("perpendicular_conduction", C.c_int32), ("viscosity", C.c_int32),
        ("viscous_heating", C.c_int32), ("hall", C.c_int32), ("electron_inertia", C.c_int32),
        ("gravity", C.c_int32), ("boundary_type", C.c_int32), ("coaxial", C.c_int32),
        ("reserved", C.c_int32),
        ("k2", C.c_double), ("k3", C.c_double), ("gamma", C.c_double),
        ("viscosity_value", C.c_double), ("electron_fraction", C.c_double),
        ("gauss_nodes", C.c_double * 4), ("gauss_weights", C.c_double * 4),
    ]


class CArnoldi(C.Structure):
    _fields_ = [
        ("nev", C.c_int32), ("ncv", C.c_int32), ("maxiter", C.c_int32),
        ("which", C.c_char * 2), ("pad", C.c_int16),
        ("tol", C.c_double), ("sigma_re", C.c_double), ("sigma_im", C.c_double),
        ("refine_steps", C.c_int32), ("reserved", C.c_int32),
    ]


class CStats(C.Structure):
    _fields_ = [
        ("info", C.c_int32), ("nconv", C.c_int32), ("n_op", C.c_int32), ("n_bx", C.c_int32),
        ("n_reorth", C.c_int32), ("n_restart", C.c_int32), ("lu_info", C.c_int32),
        ("reserved", C.c_int32),
        ("t_factor_ms", C.c_double), ("t_iter_ms", C.c_double), ("t_extract_ms", C.c_double),
    ]


_P = C.c_void_p
_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int32)

# name -> (restype, argtypes); every symbol include/legolas_b200.h declares
SIGNATURES = {
    "lgpu_create": (C.c_int, [C.POINTER(_P), C.c_int32, C.c_int32]),
    "lgpu_destroy": (C.c_int, [_P]),
    "lgpu_last_error": (C.c_char_p, [_P]),
    "lgpu_set_stream": (C.c_int, [_P, _P]),
    "lgpu_synchronize": (C.c_int, [_P]),
    "lgpu_set_sm_limit": (C.c_int, [_P, C.c_int32]),
    "lgpu_assemble": (C.c_int, [_P, C.POINTER(CSettings), _P, _P, C.POINTER(_P)]),
    "lgpu_assemble_device": (C.c_int, [_P, C.POINTER(CSettings), _P, _P, C.POINTER(_P)]),
    "lgpu_matrix_dim": (C.c_int, [_P, _IP]),
    "lgpu_export_coo": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int64), _P, _P, _P]),
    "lgpu_export_blocks": (C.c_int, [_P, C.c_int32, _P]),
    "lgpu_import_coo": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int64, _P, _P, _P]),
    "lgpu_factorize": (C.c_int, [_P, C.c_double, C.c_double, _IP]),
    "lgpu_solve": (C.c_int, [_P, _P, _P, C.c_int32]),
    "lgpu_matvec": (C.c_int, [_P, C.c_int32, _P, _P]),
    "lgpu_apply_op": (C.c_int, [_P, _P, _P, C.c_int32]),
    "lgpu_apply_op_device": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _DP]),
    "lgpu_shift_invert": (C.c_int, [_P, C.POINTER(CArnoldi), _P, _P, _P, C.POINTER(CStats)]),
    "lgpu_shift_invert_device": (C.c_int, [_P, C.POINTER(CArnoldi), _P, _P, _P, C.POINTER(CStats)]),
    "lgpu_arnoldi_general": (C.c_int, [_P, C.POINTER(CArnoldi), _P, _P, _P, C.POINTER(CStats)]),
    "lgpu_residuals": (C.c_int, [_P, C.c_int32, _P, _P, _DP]),
    "lgpu_eigenfunctions": (C.c_int, [_P, _P, C.c_int32, _IP, _P]),
    "lgpu_inverse_iteration": (C.c_int, [_P, C.c_double, C.c_double, C.c_int32, C.c_double, _P, _P,
                                         C.POINTER(CStats)]),
    "lgpu_host_alloc": (C.c_void_p, [C.c_size_t]),
    "lgpu_host_free": (None, [C.c_void_p]),
    "lgpu_zlarnv": (C.c_int, [_IP, C.c_int32, _P]),
    "lgpu_counters": (C.c_int, [_P, C.POINTER(C.c_int64), C.c_int32]),
    "lgpu_set_profiling": (C.c_int, [_P, C.c_int32]),
    "lgpu_profile_read": (C.c_int, [_P, _DP, C.POINTER(C.c_int64), _DP, C.c_int32, C.c_int32]),
    "lgpu_phase_times": (C.c_int, [_P, _DP, _DP, _DP, _DP]),
}

_lib = None


def load():
    """Load the shared library (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LgpuError(ENOGPU, f"{LIB_PATH} is missing: run `python -m legolas_b200.build` "
                                "(no CPU fallback exists)")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if a declared symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib
