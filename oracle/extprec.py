"""Loader for ``oracle/extprec.c`` (80-bit block-tridiagonal products of the arbiter).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The shared object is built on demand with
gcc into ``oracle/_build/`` (git-ignored, travels to the GPU box with the snapshot);
``__graft_entry__.build()`` builds it too.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "extprec.c")
LIB = os.path.join(HERE, "_build", "libextprec.so")
_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if out.returncode != 0:
        raise RuntimeError("gcc failed on oracle/extprec.c:\n" + out.stdout.decode())
    return LIB


def _load():
    global _lib
    if _lib is None:
        assert np.dtype(np.clongdouble).itemsize == 32, "needs x87 long double (x86-64)"
        _lib = ctypes.CDLL(build())
        _lib.bt_gemv_ld.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p]
        _lib.bt_gemv_ld.restype = None
    return _lib


def gemv_ld(A_blocks, B_blocks, sigma, x_ld, z_ld=None, sign=1):
    """sign * (A - sigma B) x + z in 80-bit arithmetic; A, B: (G, 3, d, d) complex128 blocks
    (B may be None: the matrix is A), x, z: clongdouble vectors.  Returns a clongdouble vector."""
    lib = _load()
    A_blocks = np.ascontiguousarray(A_blocks, dtype=np.complex128)
    G, _, d, _ = A_blocks.shape
    if B_blocks is not None:
        B_blocks = np.ascontiguousarray(B_blocks, dtype=np.complex128)
    x_ld = np.ascontiguousarray(x_ld, dtype=np.clongdouble)
    if z_ld is not None:
        z_ld = np.ascontiguousarray(z_ld, dtype=np.clongdouble)
    sig = np.array([complex(sigma).real, complex(sigma).imag], dtype=np.float64)
    y = np.empty(G * d, dtype=np.clongdouble)
    lib.bt_gemv_ld(G, d, A_blocks.ctypes.data, B_blocks.ctypes.data if B_blocks is not None else None,
                   sig.ctypes.data, int(sign), x_ld.ctypes.data,
                   z_ld.ctypes.data if z_ld is not None else None, y.ctypes.data)
    return y
