"""howmny = 'P' of dneupd / zneupd (dneupd.f:97-104, zneupd.f:83-90): instead of Ritz vectors the routine returns an
orthonormal basis of the converged invariant subspace (Schur vectors); the matching upper (quasi-)triangular Schur
matrix is left in workl(ipntr(12)) (dneupd.f:246-249).  Oracle and the product's host control code (CPU test double):
A Z = Z T, Z^H Z = I, eig(T) = the returned Ritz values, and both take the same path."""
import numpy as np
import pytest

from backends import HostDouble, Oracle
from problems import complex_tridiag, convdiff2d

BACKENDS = {"oracle": Oracle, "hostlogic": HostDouble}


def _schur_block(r, ncv, k):
    T = r.workl_eupd[r.ipntr_eupd[11] - 1: r.ipntr_eupd[11] - 1 + ncv * ncv].reshape(ncv, ncv).T
    return T[:k, :k]


@pytest.mark.parametrize("backend", ["oracle", "hostlogic"])
def test_dneupd_schur_vectors(backend):
    A = convdiff2d(12, rho=10.0)
    n, nev, ncv = A.shape[0], 5, 20
    r0 = np.random.default_rng(3).uniform(-1, 1, n)
    r = BACKENDS[backend]().solve(lambda x: A @ x, n, nev, ncv, "LM", sym=False, tol=1e-10, mxiter=3000, resid=r0,
                                  howmny="P")
    assert r.info == 0 and r.ierr == 0
    k = r.nconv
    Z = r.z[:k].T
    T = _schur_block(r, ncv, k)
    assert np.abs(Z.T @ Z - np.eye(k)).max() <= 1e-10
    assert np.abs(A @ Z - Z @ T).max() <= 1e-8 * np.abs(T).max()
    assert np.abs(np.tril(T, -2)).max() == 0.0                                 # quasi-triangular
    ev = np.linalg.eigvals(T)
    got = r.dr[:k] + 1j * r.di[:k]
    assert np.abs(np.sort_complex(ev) - np.sort_complex(got)).max() <= 1e-9 * np.abs(got).max()
    if backend == "hostlogic":
        o = Oracle().solve(lambda x: A @ x, n, nev, ncv, "LM", sym=False, tol=1e-10, mxiter=3000, resid=r0, howmny="P")
        assert (o.nconv, int(o.iparam[2]), int(o.iparam[8])) == (r.nconv, int(r.iparam[2]), int(r.iparam[8]))
        assert np.abs(np.abs(o.z[:k]) - np.abs(r.z[:k])).max() <= 1e-8


@pytest.mark.parametrize("backend", ["oracle", "hostlogic"])
def test_zneupd_schur_vectors(backend):
    A = complex_tridiag(150)
    n, nev, ncv = A.shape[0], 4, 18
    rng = np.random.default_rng(4)
    r0 = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    r = BACKENDS[backend]().solve_complex(lambda x: A @ x, n, nev, ncv, "LM", tol=1e-10, mxiter=3000, resid=r0,
                                          howmny="P")
    assert r.info == 0 and r.ierr == 0
    k = min(r.nconv, nev)
    Z = r.z[:k].T
    T = _schur_block(r, ncv, k)
    assert np.abs(Z.conj().T @ Z - np.eye(k)).max() <= 1e-10
    assert np.abs(A @ Z - Z @ T).max() <= 1e-8 * np.abs(T).max()
    assert np.abs(np.tril(T, -1)).max() == 0.0                                 # upper triangular
    assert np.abs(np.sort_complex(np.diag(T)) - np.sort_complex(r.d[:k])).max() <= 1e-9 * np.abs(r.d).max()
    if backend == "hostlogic":
        o = Oracle().solve_complex(lambda x: A @ x, n, nev, ncv, "LM", tol=1e-10, mxiter=3000, resid=r0, howmny="P")
        assert (o.nconv, int(o.iparam[2]), int(o.iparam[8])) == (r.nconv, int(r.iparam[2]), int(r.iparam[8]))
        assert np.abs(np.abs(o.z[:k]) - np.abs(r.z[:k])).max() <= 1e-8


@pytest.mark.parametrize("family", ["sym", "nonsym", "cplx"])
def test_zero_start_vector_exits_with_info_minus_9(family):
    """info = -9 (dsaupd.f:263, dsaup2.f:334-343, dnaup2.f:324-331, znaup2.f:326-333): a zero start vector -- given by
    the caller or produced by OP -- ends the solve at once; what iparam(3)/iparam(5) hold then differs between the
    families (label 1100 vs 1200 of *aup2) and must be what the oracle leaves there."""
    from problems import laplace2d
    n = 42
    A = {"sym": laplace2d(7, 6).toarray(), "nonsym": convdiff2d(7, 5.0).toarray()[:n, :n],
         "cplx": complex_tridiag(n).toarray()}[family]
    zero_op = np.zeros((n, n))
    for op_mat, r0 in ((A, np.zeros(n)), (zero_op, np.ones(n))):
        res = []
        for cls in (Oracle, HostDouble):
            if family == "cplx":
                r = cls().solve_complex(lambda x: (op_mat @ x).astype(complex), n, 3, 12, "LM", tol=1e-8, mxiter=50,
                                        resid=r0.astype(complex), eupd=False)
            else:
                r = cls().solve(lambda x: op_mat @ x, n, 3, 12, "LM", sym=(family == "sym"), tol=1e-8, mxiter=50,
                                resid=r0, eupd=False)
            res.append((r.info, int(r.iparam[2]), int(r.iparam[4]), int(r.iparam[8])))
        assert res[0][0] == -9 and res[0] == res[1], res
