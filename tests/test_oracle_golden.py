"""Pin the CPU oracle (oracle/libref_arpack.so) on every known-answer test the reference holds for this path
 and on the committed golden fixtures.  The reference is Fortran and cannot be compiled here, so
these analytic / LAPACK-stream answers are what anchors parity."""
import ctypes as C
import json
import os
import threading

import numpy as np
import pytest

from backends import Oracle, lib as oracle_lib
from problems import convdiff2d, dssimp_av, dssimp_exact, laplace2d

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def test_icb_arpack_c_known_answer():
    """TESTS/icb_arpack_c.c:31-91: diag(1..1000), nev=9, ncv=19, 'LM', tol=1e-6 -> 992..1000 within 1e-5."""
    n = 1000
    diag = np.arange(1, n + 1, dtype=float)
    r = Oracle().solve(lambda x: diag * x, n, 9, 19, "LM", tol=1e-6, mxiter=10000, c_abi_tol=True)
    assert r.info == 0 and r.ierr == 0 and r.iparam[4] >= 9
    assert np.abs(r.d - np.arange(992, 1001)).max() < 1e-5
    g = golden("diag1000_sym.json")
    assert [int(r.iparam[2]), int(r.iparam[4]), int(r.iparam[8]), int(r.iparam[10])] == g["counts"]
    assert np.abs(r.d - np.array(g["d"])).max() < 1e-9


def test_bug_1315_double_known_answer():
    """TESTS/bug_1315_double.c:23-84: dnaupd/dneupd, tol=0 -> dr[i] = 1000 - i within 1e-6."""
    n = 1000
    diag = np.arange(1, n + 1, dtype=float)
    r = Oracle().solve(lambda x: diag * x, n, 9, 19, "LM", sym=False, tol=0.0, mxiter=10 * n, c_abi_tol=True)
    assert r.info == 0 and r.ierr == 0
    assert np.abs(r.dr[:9] - (1000 - np.arange(9))).max() < 1e-6
    assert np.all(r.di[:9] == 0)


def test_bug_1315_single_known_answer():
    """TESTS/bug_1315_single.c:75: float twin, 1e-1."""
    n = 1000
    diag = np.arange(1, n + 1, dtype=np.float32)
    r = Oracle().solve(lambda x: diag * x, n, 9, 19, "LM", sym=False, tol=0.0, mxiter=10 * n, c_abi_tol=True,
                       dtype=np.float32)
    assert r.info == 0 and r.ierr == 0
    assert np.abs(r.dr[:9] - (1000 - np.arange(9))).max() < 1e-1


def test_dssimp_known_answer():
    """EXAMPLES/SIMPLE/dssimp.f:180-287 (BASELINE config 1): the four largest eigenvalues of the scaled 2-D
    Laplacian nx=10 are 891.1667098902916, 919.7806545598542 (x2), 948.3945992294168."""
    nx = 10
    r = Oracle().solve(dssimp_av(nx), nx * nx, 4, 20, "LM", tol=0.0, mxiter=300)
    assert r.info == 0 and r.ierr == 0 and r.nconv == 4
    exact = np.array([891.1667098902916, 919.7806545598542, 919.7806545598542, 948.3945992294168])
    assert np.abs(dssimp_exact(nx, 4) - exact).max() < 1e-10
    assert np.abs(r.d - exact).max() < 1e-10
    # the documented sample run (DOCUMENTS/debug.doc:42-45, single precision) took 8 iterations / 125 OP*x /
    # 125 re-orthogonalisation steps; the double-precision oracle lands on 8 / 126 / 125
    assert int(r.iparam[2]) == 8 and r.stats["nopx"] == 126 and r.stats["nrorth"] == 125
    for i in range(4):
        assert np.linalg.norm(dssimp_av(nx)(r.z[i]) - r.d[i] * r.z[i]) < 1e-10 * abs(r.d[i])


def test_dlarnv_stream_golden():
    """The LAPACK stream dgetv0.f:236 draws from (values measured through scipy_dlarnv_)."""
    g = golden("dlarnv.json")
    L = oracle_lib()
    for case in g["cases"]:
        seed = np.array(case["seed"], dtype=np.int32)
        x = np.zeros(len(case["values"]))
        L.ref_dlarnv2(seed.ctypes.data_as(C.POINTER(C.c_int)), len(x), x.ctypes.data_as(C.POINTER(C.c_double)))
        assert np.array_equal(x, np.array(case["values"]))
        assert list(seed) == case["seed_after"]


def test_second_solve_continues_the_random_stream():
    """CHANGES:9 / dgetv0.f:164,202-208: iseed is SAVE'd, so a second solve in the same process draws new numbers."""
    n = 200
    diag = np.arange(1, n + 1, dtype=float)
    o = Oracle()
    firsts = []
    for _ in range(2):
        seen = []

        def op(x):
            if not seen:
                seen.append(x.copy())
            return diag * x
        o.solve(op, n, 3, 10, "LM", tol=1e-8, mxiter=100)
        firsts.append(seen[0])
    assert not np.array_equal(firsts[0], firsts[1])
    seed = np.array([1, 3, 5, 7], dtype=np.int32)
    x = np.zeros(2 * n)
    oracle_lib().ref_dlarnv2(seed.ctypes.data_as(C.POINTER(C.c_int)), n, x[:n].ctypes.data_as(C.POINTER(C.c_double)))
    oracle_lib().ref_dlarnv2(seed.ctypes.data_as(C.POINTER(C.c_int)), n, x[n:].ctypes.data_as(C.POINTER(C.c_double)))
    assert np.array_equal(firsts[0], x[:n]) and np.array_equal(firsts[1], x[n:])


class LogicalRanks:
    """P logical PARPACK ranks in one process: one thread per rank, barrier-based all-reduce (no MPI in the image)."""

    def __init__(self, nranks):
        self.P = nranks
        self.barrier = threading.Barrier(nranks)
        self.slots = [None] * nranks
        self.count = 0

    def allreduce_for(self, rank):
        def ar(arr, op):
            self.slots[rank] = arr.copy()
            self.barrier.wait()
            stack = np.stack(self.slots)
            out = stack.sum(0) if op == 0 else (stack.max(0) if op == 1 else stack.min(0))
            self.barrier.wait()
            return out
        return ar

    def run(self, fn):
        res = [None] * self.P
        err = []

        def work(r):
            try:
                res[r] = fn(r, self.allreduce_for(r))
            except Exception as e:  # pragma: no cover
                err.append(e)
                self.barrier.abort()
        ts = [threading.Thread(target=work, args=(r,)) for r in range(self.P)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        if err:
            raise err[0]
        return res


def split_rows(n, nranks):
    """Row distribution of PARPACK/TESTS/MPI/icb_parpack_c.c:60-69: N/nprocs each, remainder on the first ranks."""
    base, rem = divmod(n, nranks)
    counts = [base + (1 if r < rem else 0) for r in range(nranks)]
    offs = np.concatenate([[0], np.cumsum(counts)])
    return counts, offs


@pytest.mark.parametrize("nranks", [2, 3])
def test_icb_parpack_c_known_answer(nranks):
    """PARPACK/TESTS/MPI/icb_parpack_c.c:30-102: pdsaupd/pdseupd on diag(1..1000) split over the ranks,
    d = 992..1000 on every rank within 1e-5."""
    N = 1000
    counts, offs = split_rows(N, nranks)
    world = LogicalRanks(nranks)

    def rank_main(r, ar):
        diag = np.arange(offs[r] + 1, offs[r + 1] + 1, dtype=float)
        o = Oracle(rank=r, nranks=nranks, allreduce=ar)
        return o.solve(lambda x: diag * x, counts[r], 9, 19, "LM", tol=1e-6, mxiter=10000, c_abi_tol=True)
    res = world.run(rank_main)
    for r in res:
        assert r.info == 0 and r.ierr == 0
        assert np.abs(r.d - np.arange(992, 1001)).max() < 1e-5
    # replicated host state is identical on every rank
    for r in res[1:]:
        assert np.array_equal(r.d, res[0].d) and np.array_equal(r.iparam, res[0].iparam)
    # Ritz vectors: the row blocks assemble to unit vectors of the global problem
    z = np.concatenate([r.z[:9, :] for r in res], axis=1)
    assert np.allclose(np.linalg.norm(z, axis=1), 1.0, atol=1e-10)


def test_parpack_start_vector_differs_from_serial():
    """Appendix B.11: pdgetv0 does not apply OP to the start vector for bmat='I' and seeds per rank
    (pdgetv0.f:233-245,285) -> nopx starts at 0, and rank 0 draws from seed {1,0,0,1}."""
    n = 300
    diag = np.arange(1, n + 1, dtype=float)
    seen = []

    def op(x):
        seen.append(x.copy())
        return diag * x
    o = Oracle(rank=0, nranks=1, allreduce=lambda a, op_: a)
    r = o.solve(op, n, 3, 12, "LM", tol=1e-8, mxiter=200)
    assert r.info == 0
    seed = np.array([1, 0, 0, 1], dtype=np.int32)
    x = np.zeros(n)
    oracle_lib().ref_dlarnv2(seed.ctypes.data_as(C.POINTER(C.c_int)), n, x.ctypes.data_as(C.POINTER(C.c_double)))
    assert abs(x[0] - 0.48587830215175387) < 1e-16 and abs(x[1] - 0.8467738528933708) < 1e-16
    # first hand-off is already a Lanczos step: x = r0/||r0||
    assert np.allclose(seen[0], x / np.linalg.norm(x), rtol=0, atol=1e-15)


@pytest.mark.parametrize("which", ["LA", "SA", "LM", "SM", "BE"])
def test_oracle_laplace_all_which(which):
    nx, ny, nev, ncv = 14, 11, 4, 16
    A = laplace2d(nx, ny)
    ev = np.sort(np.linalg.eigvalsh(A.toarray()))
    r = Oracle().solve(lambda x: A @ x, nx * ny, nev, ncv, which, tol=1e-12, mxiter=3000)
    assert r.info == 0 and r.ierr == 0 and r.nconv == nev
    if which in ("LA", "LM"):
        want = ev[-nev:]
    elif which == "SA":
        want = ev[:nev]
    elif which == "SM":
        want = ev[np.argsort(np.abs(ev))[:nev]]
    else:
        want = np.concatenate([ev[:nev // 2], ev[-(nev - nev // 2):]])
    assert np.abs(np.sort(r.d) - np.sort(want)).max() < 1e-9


def test_oracle_nonsym_convdiff():
    """EXAMPLES/SIMPLE/dnsimp.f-style operator (rho=100): dense eigenvalues as the answer."""
    nx = 10
    A = convdiff2d(nx, 100.0)
    ev = np.linalg.eigvals(A.toarray())
    for which, key in (("SM", lambda e: np.abs(e)), ("LR", lambda e: -e.real)):
        r = Oracle().solve(lambda x: A @ x, nx * nx, 4, 20, which, sym=False, tol=1e-10, mxiter=500)
        assert r.info == 0 and r.ierr == 0
        got = r.dr[:r.nconv] + 1j * r.di[:r.nconv]
        for lam in got:
            assert np.abs(ev - lam).min() < 1e-7 * abs(lam)
        for k in range(r.nconv):  # real eigenpairs: residual check as arpackSolver.hpp:297-352
            if r.di[k] == 0:
                assert np.linalg.norm(A @ r.z[k] - r.dr[k] * r.z[k]) < 1e-6 * abs(r.dr[k])


def test_oracle_shift_invert_and_generalized():
    """Modes 3 (shift-invert) and 2 (generalised, M^-1 A) against dense solutions (TESTS/bug_1323.f is mode 3)."""
    import scipy.linalg as sl
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla
    n = 100
    A = sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]).tocsc()
    M = sp.diags([np.ones(n - 1) / 6, 4 * np.ones(n) / 6, np.ones(n - 1) / 6], [-1, 0, 1]).tocsc()
    ev = np.sort(np.linalg.eigvalsh(A.toarray()))
    # mode 3, bmat='I', sigma=0 (bug_1323.f): OP = inv(A - sigma I)
    lu = sla.splu(A)
    r = Oracle().solve(lambda x: lu.solve(x), n, 4, 10, "LM", tol=0.0, mxiter=300, mode=3, sigma=0.0)
    assert r.info == 0 and r.ierr == 0
    assert np.abs(np.sort(r.d) - ev[:4]).max() < 1e-10
    # mode 3, bmat='G': OP = inv(A - sigma M) M
    gev = np.sort(sl.eigh(A.toarray(), M.toarray(), eigvals_only=True))
    sigma = 0.0
    lu2 = sla.splu((A - sigma * M).tocsc())

    def op3(x, is_bx=False):
        return lu2.solve(x if is_bx else M @ x)
    r = Oracle().solve(op3, n, 4, 10, "LM", tol=0.0, mxiter=300, mode=3, bmat="G", sigma=sigma, bop=lambda x: M @ x)
    assert r.info == 0 and r.ierr == 0
    assert np.abs(np.sort(r.d) - gev[:4]).max() < 1e-9
    # mode 2: OP = inv(M) A, x overwritten with A x
    luM = sla.splu(M)

    def op2(x):
        ax = A @ x
        return luM.solve(ax), ax
    r = Oracle().solve(op2, n, 4, 10, "LM", tol=0.0, mxiter=300, mode=2, bmat="G", bop=lambda x: M @ x)
    assert r.info == 0 and r.ierr == 0
    assert np.abs(np.sort(r.d) - gev[-4:]).max() < 1e-8


# ----------------------------------------------------------------------------------------------------
# Independent cross-check: SciPy >= 1.15 ships its own C translation of ARPACK-NG's [ds]saupd/[ds]naupd
# (scipy/sparse/linalg/_eigen/arpack/_arpacklib) -- a second, unrelated restatement of the same Fortran.
# With the same start vector both must take the same path: restart count, OP*x count, nconv, eigenvalues.
# ----------------------------------------------------------------------------------------------------
def _scipy_arpack(sym, A, n, nev, ncv, which, r0, tol, tp="d", maxiter=3000):
    import importlib
    mod = importlib.import_module("scipy.sparse.linalg._eigen.arpack.arpack")
    if not hasattr(mod, "_arpacklib"):
        pytest.skip("this SciPy does not ship the C translation of ARPACK")
    cls = mod._SymmetricArpackParams if sym else mod._UnsymmetricArpackParams
    P = cls(n, nev, tp, lambda x: A @ x, ncv=ncv, v0=r0.astype(np.float64 if tp == "d" else np.float32).copy(),
            maxiter=maxiter, which=which, tol=tol)
    nopx = 0
    while not P.converged:
        P.iterate()
        nopx += P.arpack_dict["ido"] in (1, 5)
    vals = P.extract(False)
    return dict(nconv=int(P.arpack_dict["nconv"]), restarts=int(P.arpack_dict["iter"]), nopx=int(nopx),
                vals=np.asarray(vals))


@pytest.mark.parametrize("which", ["LA", "SA", "LM", "SM", "BE"])
def test_oracle_follows_scipy_arpack_translation_sym(which):
    from problems import laplace2d
    A = laplace2d(17, 13)
    n = A.shape[0]
    r0 = np.random.default_rng(7).uniform(-1, 1, n)
    s = _scipy_arpack(True, A, n, 5, 18, which, r0, 1e-10)
    b = Oracle().solve(lambda x: A @ x, n, 5, 18, which, tol=1e-10, mxiter=3000, resid=r0)
    assert (s["nconv"], s["restarts"], s["nopx"]) == (int(b.nconv), int(b.iparam[2]), int(b.iparam[8]))
    assert np.abs(np.sort(s["vals"]) - np.sort(b.d)).max() < 1e-12 * np.abs(b.d).max()


@pytest.mark.parametrize("which", ["LM", "LR", "SR", "SM"])
def test_oracle_follows_scipy_arpack_translation_nonsym(which):
    from problems import convdiff2d
    A = convdiff2d(14, rho=10.0)
    n = A.shape[0]
    r0 = np.random.default_rng(7).uniform(-1, 1, n)
    s = _scipy_arpack(False, A, n, 4, 16, which, r0, 1e-10)
    b = Oracle().solve(lambda x: A @ x, n, 4, 16, which, sym=False, tol=1e-10, mxiter=3000, resid=r0)
    assert (s["nconv"], s["restarts"], s["nopx"]) == (int(b.nconv), int(b.iparam[2]), int(b.iparam[8]))
    ev = np.sort_complex(b.dr[:4] + 1j * b.di[:4])
    assert np.abs(np.sort_complex(np.asarray(s["vals"], dtype=complex))[:4] - ev).max() < 1e-10 * np.abs(ev).max()


def test_oracle_follows_scipy_arpack_translation_larger_and_float():
    from problems import laplace2d
    A = laplace2d(60, 45)
    n = A.shape[0]
    r0 = np.random.default_rng(21).uniform(-1, 1, n)
    s = _scipy_arpack(True, A, n, 8, 30, "LA", r0, 1e-12)
    b = Oracle().solve(lambda x: A @ x, n, 8, 30, "LA", tol=1e-12, mxiter=3000, resid=r0)
    assert (s["nconv"], s["restarts"], s["nopx"]) == (int(b.nconv), int(b.iparam[2]), int(b.iparam[8]))
    assert np.abs(np.sort(s["vals"]) - np.sort(b.d)).max() < 1e-12 * np.abs(b.d).max()
    A32 = A.astype(np.float32)
    s = _scipy_arpack(True, A32, n, 4, 16, "LA", r0, 1e-4, tp="f")
    b = Oracle().solve(lambda x: A32 @ x, n, 4, 16, "LA", tol=1e-4, mxiter=3000, resid=r0, dtype=np.float32)
    assert s["nconv"] == int(b.nconv) == 4
    assert np.abs(np.sort(s["vals"]) - np.sort(b.d)).max() < 1e-4 * np.abs(b.d).max()


# ---- committed golden vectors from SciPy's ARPACK translation (tests/golden/scipy_arpack_cases.json) ----
import golden_cases  # noqa: E402


@pytest.mark.parametrize("c", golden_cases.load(), ids=golden_cases.case_id)
def test_oracle_reproduces_committed_scipy_arpack_vectors(c):
    A = golden_cases.PROBLEMS[c["problem"]]()
    n = A.shape[0]
    r = Oracle().solve(lambda x: A @ x, n, c["nev"], c["ncv"], c["which"], sym=c["sym"], tol=c["tol"], mxiter=3000,
                       resid=golden_cases.start_vector(c, n))
    golden_cases.check_against_golden(c, r, c["nev"])
