"""CPU oracle for the Legolas hot path (assembly + shift-invert Arnoldi).

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (NumPy / SciPy
LAPACK + ARPACK) of the reference algorithm; it exists to *check* the CUDA
path and to time a CPU baseline next to it.  Nothing under ``legolas_b200/``
may import it: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do.

Parity pinning: the restatement is pinned by the reference's own golden data
(``tests/golden/*.npz``, produced from the reference's datfiles by
``tests/golden/make_golden.py``) and by the known-answer vectors of the
reference's pFUnit tests; see ``tests/test_oracle_*.py``.

The reference itself is Fortran 2008 and cannot be compiled in this image
(no gfortran/flang), so there is no ``oracle/_ref`` build.
"""
