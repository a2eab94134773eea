"""The run-time tracing of the reference (COMMON /debug/ set through debug_c, ICB/debug_c.h; messages of
dsaupd.f:630-680, dsaup2.f:404-413,494-504, dnaupd.f:630-680, znaupd.f:603-660, dgetv0.f:398-401): the product's host
control code prints the path messages and the exit summary when the routine's level is raised, nothing at the default
levels, and the counters in the summary are those of iparam(9:11).  Checked on the CPU test double; the product sets the
same levels through debug_c."""
import ctypes as C
import re

import numpy as np

from backends import HostDouble, hostdouble_lib
from problems import complex_tridiag, convdiff2d, laplace2d

NAMES = ["logfil", "ndigit", "mgetv0", "msaupd", "msaup2", "msaitr", "mseigt", "msapps", "msgets", "mseupd", "mnaupd",
         "mnaup2", "mnaitr", "mneigh", "mnapps", "mngets", "mneupd", "mcaupd", "mcaup2", "mcaitr", "mceigh", "mcapps",
         "mcgets", "mceupd"]


def _debug(**kw):
    v = {"logfil": 6, "ndigit": -3}
    v.update(kw)
    arr = np.array([v.get(n, 0) for n in NAMES], dtype=np.int32)
    hostdouble_lib().hd_debug(arr.ctypes.data_as(C.POINTER(C.c_int)))


def test_default_levels_print_nothing(capfd):
    _debug()
    A = laplace2d(12, 9)
    r = HostDouble().solve(lambda x: A @ x, A.shape[0], 3, 12, "LA", tol=1e-8, mxiter=300,
                           resid=np.random.default_rng(0).uniform(-1, 1, A.shape[0]))
    assert r.info == 0
    out, err = capfd.readouterr()
    assert out == "" and err == ""


def test_symmetric_trace_and_summary(capfd):
    A = laplace2d(12, 9)
    try:
        _debug(msaupd=1, msaup2=3, mgetv0=1)
        r = HostDouble().solve(lambda x: A @ x, A.shape[0], 3, 12, "LA", tol=1e-8, mxiter=300,
                               resid=np.random.default_rng(0).uniform(-1, 1, A.shape[0]), eupd=False)
    finally:
        _debug()
    out, _ = capfd.readouterr()
    assert r.info == 0
    # one "Start of major iteration" banner per restart, numbered 1..iparam(3)
    starts = re.findall(r"_saup2: \*\*\*\* Start of major iteration number \*\*\*\*\n -+\n\s+1 -\s+1:\s+(\d+)", out)
    assert [int(s) for s in starts] == list(range(1, int(r.iparam[2]) + 1))
    assert "_saup2: NEV, NP, NCONV are" in out and "_saup2: The eigenvalues of H" in out
    assert "_getv0: B-norm of initial / restarted starting vector" in out
    assert "_saupd: final Ritz values" in out and "= Symmetric implicit Arnoldi update code" in out
    m = re.search(r"Total number of OP\*x operations\s+=\s+(\d+)", out)
    assert m and int(m.group(1)) == int(r.iparam[8])
    m = re.search(r"Total number of reorthogonalization steps\s+=\s+(\d+)", out)
    assert m and int(m.group(1)) == int(r.iparam[10])
    # dvout format (1P, D12.3) with ndigit = -3: the largest Ritz value appears as d.dddD+dd
    top = np.sort(np.linalg.eigvalsh(A.toarray()))[-1]
    mant, ex = f"{top:.3E}".split("E")
    assert f"{mant}D{ex}" in out


def test_nonsymmetric_and_complex_summaries(capfd):
    A = convdiff2d(8, rho=10.0)
    Zm = complex_tridiag(60)
    rng = np.random.default_rng(3)
    try:
        _debug(mnaupd=1, mnaup2=1, mcaupd=1, mcaup2=1, ndigit=-6)
        r = HostDouble().solve(lambda x: A @ x, A.shape[0], 3, 14, "LM", sym=False, tol=1e-8, mxiter=300,
                               resid=rng.uniform(-1, 1, A.shape[0]), eupd=False)
        z = HostDouble().solve_complex(lambda x: Zm @ x, 60, 3, 14, "LM", tol=1e-8, mxiter=300,
                                       resid=rng.uniform(-1, 1, 60) + 0j, eupd=False)
    finally:
        _debug()
    out, _ = capfd.readouterr()
    assert r.info == 0 and z.info == 0
    assert "= Nonsymmetric implicit Arnoldi update code" in out and "= Complex implicit Arnoldi update code" in out
    assert "_naupd: Real part of the final Ritz values" in out and "_naupd: The final Ritz values" in out
    its = re.findall(r"Total number update iterations\s+=\s+(\d+)", out)
    assert [int(x) for x in its] == [int(r.iparam[2]), int(z.iparam[2])]
    assert out.count("_naup2: **** Start of major iteration number ****") == int(r.iparam[2]) + int(z.iparam[2])
