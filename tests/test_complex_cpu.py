"""CPU tests of the complex Arnoldi path:
  * the oracle (oracle/ref_impl_complex.inc) against the reference's own known-answer test for znaupd_c/zneupd_c
    (TESTS/icb_arpack_c.c:98-165) and against the committed golden vectors made by SciPy's independent C translation
    of znaupd/zneupd (tests/golden/scipy_arpack_complex_cases.json);
  * the product's host control code (arpack-ng_b200/csrc/irl_complex.hpp) over the plain-loop test double against the
    oracle and the same golden vectors.
No GPU, no compute call into libarpack_b200.so."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import golden_cases
from backends import HostDouble, Oracle
from problems import complex_tridiag

BACKENDS = {"oracle": Oracle, "hostlogic": HostDouble}


def _counts(r):
    return int(r.nconv), int(r.iparam[2]), int(r.iparam[8]), int(r.iparam[10])


@pytest.mark.parametrize("backend", ["oracle", "hostlogic"])
@pytest.mark.parametrize("dtype,tol_check", [(np.complex128, 1e-5), (np.complex64, 1e-1)])
def test_icb_arpack_c_zn(backend, dtype, tol_check):
    """TESTS/icb_arpack_c.c:98-165: A = diag((i+1)(1+i)), nev=9, ncv=19, 'LM', tol=1e-6, rvec=0, random start ->
    d[i] = (992+i)(1+i) ascending, abs err <= 1e-5 per component (the float twin is ours, with a float tolerance)."""
    n, nev, ncv = 1000, 9, 19
    diag = (np.arange(1, n + 1) * (1 + 1j)).astype(dtype)
    r = BACKENDS[backend]().solve_complex(lambda x: diag * x, n, nev, ncv, "LM", tol=1e-6, mxiter=10 * n, rvec=False,
                                          dtype=dtype, c_abi_tol=True)
    assert r.info == 0 and r.ierr == 0 and r.nconv >= nev
    want = (n - (nev - 1) + np.arange(nev)) * (1 + 1j)
    assert np.abs(r.d.real - want.real).max() <= tol_check and np.abs(r.d.imag - want.imag).max() <= tol_check


@pytest.mark.parametrize("backend", ["oracle", "hostlogic"])
@pytest.mark.parametrize("c", golden_cases.load_complex(), ids=golden_cases.case_id)
def test_reproduces_committed_scipy_znaupd_vectors(backend, c):
    """Same nconv, restart count, OP*x count and eigenvalues (1e-10) as SciPy's translation of znaupd/zneupd."""
    A = golden_cases.ZPROBLEMS[c["problem"]]()
    n = A.shape[0]
    r = BACKENDS[backend]().solve_complex(lambda x: A @ x, n, c["nev"], c["ncv"], c["which"], tol=c["tol"], mxiter=3000,
                                          resid=golden_cases.start_vector_complex(c, n))
    golden_cases.check_against_golden_complex(c, r)
    Z = r.z.T
    assert (np.linalg.norm(A @ Z - Z * r.d[None, :], axis=0) <= 1e-8 * np.abs(r.d).max()).all()


@pytest.mark.parametrize("c", golden_cases.load_complex(), ids=golden_cases.case_id)
def test_hostlogic_follows_oracle_exactly(c):
    """Plain loops in the same order as the oracle's BLAS: identical path including the re-orthogonalisation count,
    eigenvectors equal up to rounding."""
    A = golden_cases.ZPROBLEMS[c["problem"]]()
    n = A.shape[0]
    r0 = golden_cases.start_vector_complex(c, n)
    a = Oracle().solve_complex(lambda x: A @ x, n, c["nev"], c["ncv"], c["which"], tol=c["tol"], mxiter=3000, resid=r0)
    b = HostDouble().solve_complex(lambda x: A @ x, n, c["nev"], c["ncv"], c["which"], tol=c["tol"], mxiter=3000,
                                   resid=r0)
    assert _counts(a) == _counts(b)
    assert np.abs(a.d - b.d).max() <= 1e-11 * np.abs(a.d).max()
    assert np.abs(np.abs(a.z) - np.abs(b.z)).max() <= 1e-8


@pytest.mark.parametrize("backend", ["oracle", "hostlogic"])
def test_random_start_is_the_zlarnv_stream(backend):
    """info = 0: the start vector comes from LAPACK zlarnv(idist=2, iseed={1,3,5,7}) (zgetv0.f:196-230); the first
    hand-off (ido = -1) exposes it as workd(ipntr(1)).  Both implementations draw the same stream, so the whole solve
    takes the same path."""
    A = complex_tridiag(120)
    n = A.shape[0]
    a = Oracle().solve_complex(lambda x: A @ x, n, 3, 14, "LM", tol=1e-9, mxiter=2000)
    b = BACKENDS[backend]().solve_complex(lambda x: A @ x, n, 3, 14, "LM", tol=1e-9, mxiter=2000)
    assert a.info == b.info == 0
    assert _counts(a)[:3] == _counts(b)[:3]
    assert np.abs(a.d - b.d).max() <= 1e-10 * np.abs(a.d).max()


@pytest.mark.parametrize("backend", ["oracle", "hostlogic"])
def test_shift_invert_mode3_and_generalized(backend):
    """Mode 3 with bmat='I' (zndrv2-style) and with bmat='G' (zndrv4-style: OP = inv(A - sigma M) M, B = M):
    eigenvalues nearest sigma against a dense solve; exercises ido = 2 hand-offs, the B-inner products, the
    back-transform and the zgeru purification (zneupd.f:820-868)."""
    A = complex_tridiag(150)
    n, nev, ncv = A.shape[0], 4, 20
    sigma = 30000.0 + 15.0j
    rng = np.random.default_rng(7)
    r0 = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    # bmat = 'I'
    lu = spla.splu((A - sigma * sp.eye(n)).tocsc())
    r = BACKENDS[backend]().solve_complex(lambda x: lu.solve(np.ascontiguousarray(x)), n, nev, ncv, "LM", tol=1e-10,
                                          mxiter=2000, resid=r0, mode=3, sigma=sigma)
    assert r.info == 0 and r.ierr == 0 and r.nconv == nev
    dense = np.linalg.eigvals(A.toarray())
    want = dense[np.argsort(np.abs(dense - sigma))[:nev]]
    assert np.abs(np.sort_complex(r.d) - np.sort_complex(want)).max() <= 1e-9 * np.abs(want).max()
    Z = r.z.T
    assert (np.linalg.norm(A @ Z - Z * r.d[None, :], axis=0) <= 1e-7 * np.abs(r.d).max()).all()
    # bmat = 'G' with a Hermitian positive definite mass matrix M
    M = sp.diags([np.full(n - 1, 1.0), np.full(n, 4.0), np.full(n - 1, 1.0)], [-1, 0, 1]).tocsr().astype(complex) / 6.0
    lug = spla.splu((A - sigma * M).tocsc())

    def op(x, bx):   # ido = -1: y = inv(A - sigma M) M x ; ido = 1: M x is supplied in workd(ipntr(3))
        return lug.solve(np.ascontiguousarray(M @ x if bx is None else bx))
    g = BACKENDS[backend]().solve_complex(op, n, nev, ncv, "LM", tol=1e-10, mxiter=2000, resid=r0, mode=3, sigma=sigma,
                                          bmat="G", bop=lambda x: M @ x)
    assert g.info == 0 and g.ierr == 0 and g.nconv == nev
    import scipy.linalg as sla
    gd = sla.eigvals(A.toarray(), M.toarray())
    gw = gd[np.argsort(np.abs(gd - sigma))[:nev]]
    assert np.abs(np.sort_complex(g.d) - np.sort_complex(gw)).max() <= 1e-9 * np.abs(gw).max()
    Zg = g.z.T
    assert (np.linalg.norm(A @ Zg - (M @ Zg) * g.d[None, :], axis=0) <= 1e-7 * np.abs(g.d).max()).all()
    assert int(g.iparam[9]) > 0   # nbx: B*x hand-offs happened


@pytest.mark.parametrize("backend", ["oracle", "hostlogic"])
def test_argument_errors_and_budget_exit(backend):
    """znaupd.f:462-489 error codes; info = 1 when the restart budget is exhausted; zneupd's stricter ncv check."""
    A = complex_tridiag(60)
    n = A.shape[0]
    B = BACKENDS[backend]
    op = lambda x: A @ x  # noqa: E731
    assert B().solve_complex(op, n, 0, 12, "LM", eupd=False).info == -2
    assert B().solve_complex(op, n, 3, 3, "LM", eupd=False).info == -3
    assert B().solve_complex(op, n, 3, n + 1, "LM", eupd=False).info == -3
    assert B().solve_complex(op, n, 3, 12, "LA", eupd=False).info == -5
    assert B().solve_complex(op, n, 3, 12, "LM", bmat="X", eupd=False).info == -6
    assert B().solve_complex(op, n, 3, 12, "LM", mode=4, eupd=False).info == -10
    assert B().solve_complex(op, n, 3, 12, "LM", bmat="G", eupd=False).info == -11
    rng = np.random.default_rng(1)
    r0 = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    a = Oracle().solve_complex(op, n, 4, 12, "SM", tol=1e-14, mxiter=2, resid=r0, eupd=False)
    b = B().solve_complex(op, n, 4, 12, "SM", tol=1e-14, mxiter=2, resid=r0, eupd=False)
    assert a.info == b.info == 1 and _counts(a)[:3] == _counts(b)[:3]
    # ncv = nev + 1 passes znaupd (ncv > nev) but zneupd answers -3 (zneupd.f:360)
    r = B().solve_complex(op, n, 3, 4, "LM", tol=1e-6, mxiter=500, resid=r0)
    assert r.info in (0, 1) and (r.get("ierr", -3) == -3 or r.nconv == 0)


def test_user_supplied_shifts_ido3():
    """ishift = 0: znaupd returns ido = 3 and applies the iparam(8) shifts the caller stores at workl(ipntr(14))
    (znaup2.f:661-685).  Exact shifts supplied by hand (the np leading entries of the sorted Ritz array, i.e. the unwanted
    ones) must reproduce the wanted eigenvalues, and the host logic must follow the oracle count for count."""
    A = complex_tridiag(100)
    n, nev, ncv = 100, 3, 14
    rng = np.random.default_rng(12)
    r0 = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)

    def shifts(ritz, bounds, npsh):
        return ritz[:npsh]          # zngets left the unwanted values first
    a = Oracle().solve_complex(lambda x: A @ x, n, nev, ncv, "LM", tol=1e-10, mxiter=500, resid=r0, ishift=0,
                               shifts=shifts)
    b = HostDouble().solve_complex(lambda x: A @ x, n, nev, ncv, "LM", tol=1e-10, mxiter=500, resid=r0, ishift=0,
                                   shifts=shifts)
    assert a.info == b.info == 0 and a.ierr == b.ierr == 0
    assert a.nshift_calls == b.nshift_calls > 0
    assert _counts(a) == _counts(b)
    assert np.abs(a.d - b.d).max() <= 1e-11 * np.abs(a.d).max()
    dense = np.linalg.eigvals(A.toarray())
    want = dense[np.argsort(-np.abs(dense))[:nev]]
    assert np.abs(np.sort_complex(a.d) - np.sort_complex(want)).max() <= 1e-9 * np.abs(want).max()


@pytest.mark.parametrize("backend", ["oracle", "hostlogic"])
def test_generalized_mode2(backend):
    """Mode 2 (znaupd.f:100-104): OP = inv(M) A, B = M with a Hermitian positive definite M -- every inner product is
    a B-inner product obtained through ido = 2 hand-offs.  Largest-magnitude eigenvalues of the pencil (A, M)."""
    import scipy.linalg as sla
    A = complex_tridiag(120)
    n, nev, ncv = 120, 3, 16
    off = (1.0 + 0.5j) * np.ones(n - 1) / 6.0
    M = sp.diags([np.conj(off), np.full(n, 4.0 / 6.0), off], [-1, 0, 1]).tocsc().astype(complex)   # Hermitian, SPD
    lu = spla.splu(M)
    rng = np.random.default_rng(21)
    r0 = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    r = BACKENDS[backend]().solve_complex(lambda x: lu.solve(A @ x), n, nev, ncv, "LM", tol=1e-10, mxiter=3000, resid=r0,
                                          mode=2, bmat="G", bop=lambda x: M @ x)
    assert r.info == 0 and r.ierr == 0 and r.nconv == nev and int(r.iparam[9]) > 0
    gd = sla.eigvals(A.toarray(), M.toarray())
    want = gd[np.argsort(-np.abs(gd))[:nev]]
    assert np.abs(np.sort_complex(r.d) - np.sort_complex(want)).max() <= 1e-9 * np.abs(want).max()
    Z = r.z.T
    assert (np.linalg.norm(A @ Z - (M @ Z) * r.d[None, :], axis=0) <= 1e-7 * np.abs(r.d).max()).all()
    # B-orthonormal Ritz vectors: z^H M z = 1
    assert np.abs(np.einsum("in,in->n", Z.conj(), M @ Z) - 1.0).max() <= 1e-8
    if backend == "hostlogic":
        o = Oracle().solve_complex(lambda x: lu.solve(A @ x), n, nev, ncv, "LM", tol=1e-10, mxiter=3000, resid=r0,
                                   mode=2, bmat="G", bop=lambda x: M @ x)
        assert _counts(o) == _counts(r) and int(o.iparam[9]) == int(r.iparam[9])


# ---------------------------------------------------------------------------------------------------------------
# PARPACK twins pznaupd/pzneupd on P logical ranks (no MPI in the image: threads + a barrier all-reduce)
# ---------------------------------------------------------------------------------------------------------------
from test_oracle_golden import LogicalRanks, split_rows  # noqa: E402


@pytest.mark.parametrize("backend", ["oracle", "hostlogic"])
@pytest.mark.parametrize("nranks", [2, 3])
def test_icb_parpack_c_zn_known_answer(backend, nranks):
    """PARPACK/TESTS/MPI/icb_parpack_c.c:104-190: pznaupd_c/pzneupd_c on diag((i+1)(1+i)), i < 1000, rows split over
    the ranks, nev=9, ncv=19, 'LM', tol=1e-6, rvec=0 -> d[i] = (992+i)(1+i) on every rank within 1e-5."""
    N, nev, ncv = 1000, 9, 19
    cnts, offs = split_rows(N, nranks)
    world = LogicalRanks(nranks)

    def rank_main(r, ar):
        diag = np.arange(offs[r] + 1, offs[r + 1] + 1) * (1 + 1j)
        return BACKENDS[backend](rank=r, nranks=nranks, allreduce=ar).solve_complex(
            lambda x: diag * x, cnts[r], nev, ncv, "LM", tol=1e-6, mxiter=10 * N, rvec=False, c_abi_tol=True)
    res = world.run(rank_main)
    want = (N - (nev - 1) + np.arange(nev)) * (1 + 1j)
    for r in res:
        assert r.info == 0 and r.ierr == 0 and r.nconv >= nev
        assert np.abs(r.d.real - want.real).max() <= 1e-5 and np.abs(r.d.imag - want.imag).max() <= 1e-5
    assert all(_counts(r) == _counts(res[0]) for r in res)       # replicated host state: identical on every rank


@pytest.mark.parametrize("nranks", [2, 3])
def test_parpack_complex_semantics_match_oracle(nranks):
    """pznaupd/pzneupd path of the host logic (per-rank zlarnv seeds, no initial OP*x for bmat = 'I', all-reduced
    coefficients and norms, eps23 with a REAL exponent) against the oracle's PARPACK mode on P logical ranks: same
    path, same eigenvalues, eigenvector rows of the right operator."""
    A = complex_tridiag(180)
    n, nev, ncv = A.shape[0], 4, 18
    cnts, offs = split_rows(n, nranks)

    def run(cls, use_r0):
        world = LogicalRanks(nranks)
        rng = np.random.default_rng(17)
        r0 = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)

        def rank_main(r, ar):
            def op(x):
                # "halo exchange": gather the full vector through the all-reduce callback (re and im parts)
                full = np.zeros(2 * n)
                full[2 * offs[r]:2 * offs[r + 1]] = x.view(np.float64)
                full = ar(full, 0).view(np.complex128)
                return (A @ full)[offs[r]:offs[r + 1]]
            return cls(rank=r, nranks=nranks, allreduce=ar).solve_complex(
                op, cnts[r], nev, ncv, "LM", tol=1e-10, mxiter=3000, resid=r0[offs[r]:offs[r + 1]] if use_r0 else None)
        return world.run(rank_main)
    for use_r0 in (True, False):
        a, b = run(HostDouble, use_r0), run(Oracle, use_r0)
        for ra, rb in zip(a, b):
            assert ra.info == rb.info == 0 and ra.ierr == rb.ierr == 0
            assert _counts(ra) == _counts(rb)
            assert np.abs(ra.d - rb.d).max() <= 1e-10 * np.abs(rb.d).max()
        # the local eigenvector blocks assemble to eigenvectors of A
        Z = np.concatenate([ra.z for ra in a], axis=1).T          # n x nev
        assert (np.linalg.norm(A @ Z - Z * a[0].d[None, :], axis=0) <= 1e-7 * np.abs(a[0].d).max()).all()
        dense = np.linalg.eigvals(A.toarray())
        want = dense[np.argsort(-np.abs(dense))[:nev]]
        assert np.abs(np.sort_complex(a[0].d) - np.sort_complex(want)).max() <= 1e-9 * np.abs(want).max()
