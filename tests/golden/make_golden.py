"""Generates the small golden fixtures under tests/golden/.

The reference (Fortran) cannot be executed in this image, so the fixtures hold
  * dlarnv.json      -- the LAPACK dlarnv(idist=2) stream the reference's dgetv0/pdgetv0 draw from, produced by the
                        un-vendored dependency itself (OpenBLAS' LAPACK through scipy_dlarnv_), with the seeds of
                        SRC/dgetv0.f:202-208 and PARPACK/SRC/MPI/pdgetv0.f:233-245;
  * diag1000_sym.json-- the answer of TESTS/icb_arpack_c.c (992..1000) together with the counts the ORACLE produces
                        on it (oracle-generated regression values, not reference-generated).
Run:  python tests/golden/make_golden.py
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from backends import Oracle, lib  # noqa: E402


def dlarnv(seed, n):
    s = np.array(seed, dtype=np.int32)
    x = np.zeros(n)
    lib().ref_dlarnv2(s.ctypes.data_as(C.POINTER(C.c_int)), n, x.ctypes.data_as(C.POINTER(C.c_double)))
    return x, [int(v) for v in s]


cases = []
for seed, n in (([1, 3, 5, 7], 6), ([1, 0, 0, 1], 6), ([1, 0, 0, 3], 6), ([1, 3, 5, 7], 200)):
    x, after = dlarnv(seed, n)
    cases.append({"seed": seed, "values": [float(v) for v in x], "seed_after": after})
json.dump({"source": "scipy_openblas dlarnv (idist=2)", "cases": cases}, open(os.path.join(HERE, "dlarnv.json"), "w"),
          indent=1)

n = 1000
diag = np.arange(1, n + 1, dtype=float)
r = Oracle().solve(lambda x: diag * x, n, 9, 19, "LM", tol=1e-6, mxiter=10000, c_abi_tol=True)
json.dump({"source": "TESTS/icb_arpack_c.c answer 992..1000; counts = oracle run (restarts, nconv, nopx, nrorth)",
           "d": [float(v) for v in r.d], "counts": [int(r.iparam[2]), int(r.iparam[4]), int(r.iparam[8]),
                                                     int(r.iparam[10])]},
          open(os.path.join(HERE, "diag1000_sym.json"), "w"), indent=1)
print("golden fixtures written")
