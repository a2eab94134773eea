"""Randomised agreement sweep (CPU): the product's host control code over the plain-loop test double against the
oracle on small random dense-ish operators -- random n, nev, ncv, `which`, tol, restart budget, with and without a
caller-supplied start vector -- for the symmetric, nonsymmetric and complex families.  Whenever the oracle needs at
most 25 restarts the two must take the SAME path (info, nconv, restarts, OP*x) and return the same eigenvalues; on
longer runs rounding differences between BLAS and plain loops legitimately accumulate into different paths (see
DESIGN.md §3), so only the short ones are compared exactly.  Seeds are fixed: the sweep is deterministic."""
import numpy as np
import pytest
import scipy.sparse as sp

from backends import HostDouble, Oracle


def _cnt(r):
    return r.info, r.nconv, int(r.iparam[2]), int(r.iparam[8])


def _case(rng, fam):
    n = int(rng.integers(6, 60))
    nev = int(rng.integers(1, max(2, min(8, n // 3))))
    lo = nev + 1 if fam == "sym" else nev + 2
    ncv = int(rng.integers(lo, min(n, lo + 14) + 1))
    rs = lambda: int(rng.integers(1 << 30))  # noqa: E731
    if fam == "sym":
        M = sp.random(n, n, density=0.3, random_state=rs()).toarray()
        A = M + M.T + np.diag(rng.uniform(-2, 2, n))
        which = str(rng.choice(["LA", "SA", "LM", "BE"]))
        r0 = rng.uniform(-1, 1, n)
    elif fam == "nonsym":
        A = sp.random(n, n, density=0.3, random_state=rs()).toarray() + np.diag(rng.uniform(-2, 2, n))
        which = str(rng.choice(["LM", "LR", "SR", "LI"]))
        r0 = rng.uniform(-1, 1, n)
    else:
        A = sp.random(n, n, density=0.3, random_state=rs()).toarray() + \
            1j * sp.random(n, n, density=0.3, random_state=rs()).toarray() + \
            np.diag(rng.uniform(-2, 2, n) + 1j * rng.uniform(-2, 2, n))
        which = str(rng.choice(["LM", "LR", "SR", "LI", "SI"]))
        r0 = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    tol = float(rng.choice([0.0, 1e-8, 1e-12]))
    mx = int(rng.choice([3, 50, 300]))
    return A, n, nev, ncv, which, tol, mx, (r0 if rng.integers(0, 2) else None)


@pytest.mark.parametrize("fam,seed", [("sym", 11), ("nonsym", 12), ("cplx", 13)])
def test_random_small_problems_take_the_oracles_path(fam, seed):
    rng = np.random.default_rng(seed)
    compared = 0
    for _ in range(40):
        A, n, nev, ncv, which, tol, mx, r0 = _case(rng, fam)
        op = lambda x, A=A: A @ x  # noqa: E731
        if fam == "cplx":
            a = Oracle().solve_complex(op, n, nev, ncv, which, tol=tol, mxiter=mx, resid=r0)
            b = HostDouble().solve_complex(op, n, nev, ncv, which, tol=tol, mxiter=mx, resid=r0)
        else:
            a = Oracle().solve(op, n, nev, ncv, which, sym=(fam == "sym"), tol=tol, mxiter=mx, resid=r0)
            b = HostDouble().solve(op, n, nev, ncv, which, sym=(fam == "sym"), tol=tol, mxiter=mx, resid=r0)
        if int(a.iparam[2]) > 25:
            continue
        compared += 1
        ctx = (fam, n, nev, ncv, which, tol, mx, r0 is not None)
        assert _cnt(a) == _cnt(b), ctx
        assert a.get("ierr", 0) == b.get("ierr", 0), ctx
        if a.info < 0 or a.get("ierr", 0) != 0 or a.nconv == 0:
            continue
        k = a.nconv if fam != "cplx" else min(a.nconv, nev)
        if fam == "sym":
            da, db = np.sort(a.d[:k]), np.sort(b.d[:k])
        elif fam == "nonsym":
            da, db = a.dr[:k] + 1j * a.di[:k], b.dr[:k] + 1j * b.di[:k]
        else:
            da, db = a.d[:k], b.d[:k]
        # same order is expected (same sorts); compare element-wise
        assert np.abs(np.asarray(da) - np.asarray(db)).max() <= 1e-8 * max(1.0, np.abs(da).max()), ctx
    assert compared >= 20
