"""The reference's regression programs (TESTS/bug_142.f, bug_58_double.f, bug_79_double_complex.f,
bug_1323.f), restated for the oracle and for the product's host control code (CPU) and, marked gpu, for the CUDA path
through the C-ABI.  The Fortran programs only `stop 1` on a defect; the assertions below are those defects plus the
answers a dense solver gives."""
import numpy as np
import pytest
import scipy.linalg as sla

from backends import HostDouble, Oracle

BACKENDS = {"oracle": Oracle, "hostlogic": HostDouble}


def bug_142_matrix():
    """TESTS/bug_142.f:118-131: an 11 x 11 column-stochastic 'Google' matrix of rank 2."""
    n = 11
    a = np.full((n, n), 0.15 / 11)
    a[0, 1:] += 0.85
    a[1:, 0] = (1 - a[0, 0]) / 10
    return a


def bug_58_matrices():
    """TESTS/bug_58_double.f:112-135: A = I except a(1,1) = 0, a(1,n) = 1 (n = 9); C = A - sigma I, sigma = -1."""
    n = 9
    a = np.eye(n)
    a[0, 0] = 0.0
    a[0, n - 1] = 1.0
    return a, a + np.eye(n)


def _check_bug_142(r, a):
    assert r.info >= 0, r.info                      # bug_142.f:170-175 `stop 1`
    assert r.ierr == 0 and r.nconv >= 1
    lam = complex(r.dr[0], r.di[0])
    assert abs(lam - 1.0) < 1e-10                   # Perron root of a column-stochastic matrix
    z = np.asarray(r.z).reshape(-1, a.shape[0])[0]
    assert np.linalg.norm(a @ z - z) < 1e-10 and np.isfinite(z).all()


@pytest.mark.parametrize("backend", ["oracle", "hostlogic"])
def test_bug_142_invariant_subspace_restart(backend):
    """ncv = n = 11 on a rank-2 matrix: the Arnoldi factorisation breaks down after two steps, so dnaitr must generate
    restart vectors (dgetv0 with j > 1) -- and must NOT force them into the range of OP when itry > 1 (the defect of
    issue 142), otherwise no new direction exists and dnaupd ends with info = -9999."""
    a = bug_142_matrix()
    r = BACKENDS[backend]().solve(lambda x: a @ x, 11, 1, 11, "LM", sym=False, tol=0.0, mxiter=300)
    _check_bug_142(r, a)
    assert r.stats["nrstrt"] > 0                    # the restart path was actually taken


def test_bug_142_hostlogic_agrees_with_oracle():
    """After the breakdown the residual norms are rounding noise (1e-31 or exactly 0 depending on the summation order
    of the BLAS), so how many of the remaining steps take the restart branch is not comparable between two
    implementations; the outcome is: one restart sweep, the same converged value."""
    a = bug_142_matrix()
    o = Oracle().solve(lambda x: a @ x, 11, 1, 11, "LM", sym=False, tol=0.0, mxiter=300)
    h = HostDouble().solve(lambda x: a @ x, 11, 1, 11, "LM", sym=False, tol=0.0, mxiter=300)
    assert (o.info, o.nconv, int(o.iparam[2])) == (h.info, h.nconv, int(h.iparam[2]))
    assert abs(o.dr[0] - h.dr[0]) < 1e-12 and o.di[0] == h.di[0] == 0.0
    assert o.stats["nrstrt"] > 0 and h.stats["nrstrt"] > 0


def _check_bug_58(r, a):
    assert r.info >= 0 and r.ierr == 0
    nconv = r.nconv
    d = r.dr[:nconv] + 1j * r.di[:nconv]
    z = np.asarray(r.z).reshape(-1, a.shape[0])[:nconv]
    assert np.isfinite(d).all() and np.isfinite(z).all()       # bug_58_double.f:415-421: NaN after the purification
    ev = np.linalg.eigvals(a)
    for lam in d:
        assert np.abs(ev - lam).min() < 1e-8                    # spectrum of A is {0, 1}
    for j in range(nconv):
        if r.di[j] == 0.0:
            assert np.linalg.norm(a @ z[j] - r.dr[j] * z[j]) <= 1e-8 * max(1.0, abs(r.dr[j]))


@pytest.mark.parametrize("backend", ["oracle", "hostlogic"])
def test_bug_58_shift_invert_purification_stays_finite(backend):
    """dnaupd mode 3, sigma = -1, nev = 4, ncv = 8 on a 9 x 9 matrix with a defective zero eigenvalue: the eigenvector
    purification of dneupd (dneupd.f:1017-1059) divides by the Ritz values of OP; the fix of issue 58 guards the zero
    ones.  Nothing returned may be NaN/Inf and the values must be eigenvalues of A."""
    a, c = bug_58_matrices()
    lu = sla.lu_factor(c)
    r = BACKENDS[backend]().solve(lambda x: sla.lu_solve(lu, x), 9, 4, 8, "LM", sym=False, tol=0.0, mxiter=300, mode=3,
                                  sigma=-1.0)
    _check_bug_58(r, a)


def _znd_av(nx):
    """OP of TESTS/bug_79_double_complex.f:232-291 (zndrv1's av/tv): 2-D convection-diffusion, rho = 100, complex."""
    import scipy.sparse as sp
    h = 1.0 / (nx + 1)
    dd, dl, du = 4.0 / h ** 2, -1.0 / h ** 2 - 0.5 * 100.0 / h, -1.0 / h ** 2 + 0.5 * 100.0 / h
    T = sp.diags([dl * np.ones(nx - 1), dd * np.ones(nx), du * np.ones(nx - 1)], [-1, 0, 1])
    off = sp.diags([np.ones(nx - 1), np.ones(nx - 1)], [-1, 1])
    return (sp.kron(sp.eye(nx), T) + sp.kron(off, -sp.eye(nx) / h ** 2)).tocsr().astype(np.complex128)


@pytest.mark.parametrize("backend", ["oracle", "hostlogic"])
def test_bug_79_first_operand_is_the_callers_resid(backend):
    """info = 1, resid = (1, 0): at the first hand-off (ido = -1) the operand workd(ipntr(1)) must be exactly the
    caller's start vector -- the program compares ||A*resid|| with ||A*workd(ipntr(1))|| for equality
    (bug_79_double_complex.f:186-226)."""
    nx = 10
    n = nx * nx
    A = _znd_av(nx)
    seen = []

    def op(x):
        if not seen:
            seen.append(x.copy())
        return A @ x
    r0 = np.ones(n, dtype=complex)
    BACKENDS[backend]().solve_complex(op, n, 4, 20, "LM", tol=0.0, mxiter=1, resid=r0, eupd=False)
    assert np.array_equal(seen[0], r0)
    assert np.linalg.norm(A @ r0) == np.linalg.norm(A @ seen[0])


@pytest.mark.parametrize("backend", ["oracle", "hostlogic"])
def test_bug_1323_mode3_without_vectors(backend):
    """TESTS/bug_1323.f: dsaupd mode 3 (sigma = 0) on the 1-D Laplacian n = 100, nev = 4, ncv = 10, 'LM', then dseupd
    with rvec = .false.: the smallest eigenvalues must come back without touching Z."""
    n = 100
    h2 = 1.0 / (n + 1) ** 2
    T = (np.diag(2 * np.ones(n)) - np.diag(np.ones(n - 1), 1) - np.diag(np.ones(n - 1), -1)) / h2
    cho = sla.cho_factor(T)
    r = BACKENDS[backend]().solve(lambda x: sla.cho_solve(cho, x), n, 4, 10, "LM", tol=0.0, mxiter=300, mode=3, sigma=0.0,
                                  rvec=False)
    assert r.info == 0 and r.ierr == 0 and r.nconv == 4
    want = np.sort(np.linalg.eigvalsh(T))[:4]
    assert np.abs(np.sort(r.d) - want).max() <= 1e-9 * want.max()


# ---------------------------------------------------------------------------------------------------------------
# the same programs on the device, through the C-ABI (host arrays and a CPU OP, as the Fortran programs have them)
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ab():
    import arpack_ng_b200 as m
    m.lib()
    return m


@pytest.mark.gpu
def test_bug_142_on_gpu(ab):
    a = bug_142_matrix()

    def op(x, y, *_):
        y[:] = a @ x
    ab.lib().ab200_reset_seed()
    r = ab.solve(op, 11, 1, 11, "LM", sym=False, tol=0.0, mxiter=300, host_buffers=True)
    _check_bug_142(r, a)
    o = Oracle().solve(lambda x: a @ x, 11, 1, 11, "LM", sym=False, tol=0.0, mxiter=300, c_abi_tol=True)
    assert (r.info, r.nconv) == (o.info, o.nconv)


@pytest.mark.gpu
def test_bug_58_on_gpu(ab):
    a, c = bug_58_matrices()
    lu = sla.lu_factor(c)

    def op(x, y, *_):
        y[:] = sla.lu_solve(lu, np.ascontiguousarray(x))
    ab.lib().ab200_reset_seed()
    r = ab.solve(op, 9, 4, 8, "LM", sym=False, tol=0.0, mxiter=300, mode=3, sigma=-1.0, host_buffers=True)
    _check_bug_58(r, a)


@pytest.mark.gpu
def test_bug_79_on_gpu(ab):
    nx = 10
    n = nx * nx
    A = _znd_av(nx)
    seen = []

    def op(x, y, *_):
        if not seen:
            seen.append(np.array(x))
        y[:] = A @ x
    r0 = np.ones(n, dtype=complex)
    ab.solve_complex(op, n, 4, 20, "LM", tol=0.0, mxiter=1, resid=r0, eupd=False, host_buffers=True)
    assert np.array_equal(seen[0], r0)


@pytest.mark.gpu
def test_bug_1323_on_gpu(ab):
    n = 100
    h2 = 1.0 / (n + 1) ** 2
    T = (np.diag(2 * np.ones(n)) - np.diag(np.ones(n - 1), 1) - np.diag(np.ones(n - 1), -1)) / h2
    cho = sla.cho_factor(T)

    def op(x, y, *_):
        y[:] = sla.cho_solve(cho, np.ascontiguousarray(x))
    ab.lib().ab200_reset_seed()
    r = ab.solve(op, n, 4, 10, "LM", tol=0.0, mxiter=300, mode=3, sigma=0.0, rvec=False, host_buffers=True)
    assert r.info == 0 and r.ierr == 0 and r.nconv == 4
    want = np.sort(np.linalg.eigvalsh(T))[:4]
    assert np.abs(np.sort(r.d) - want).max() <= 1e-9 * want.max()
