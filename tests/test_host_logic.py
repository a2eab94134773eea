"""CPU tests of the product's HOST logic: arpack-ng_b200/csrc/irl_*.hpp (state machine, ncv-sized math, error
codes, PARPACK semantics) driven through the plain-loop VecOps test double (tests/hostdouble) and compared with the
oracle.  No CUDA kernel runs here; the kernels are covered by the -m gpu tests."""
import numpy as np
import pytest

from backends import HostDouble, Oracle
from problems import convdiff2d, dssimp_av, dssimp_exact, laplace2d
from test_oracle_golden import LogicalRanks, split_rows


def counts(r):
    return int(r.nconv), int(r.iparam[2]), int(r.iparam[8]), int(r.iparam[9]), int(r.iparam[10]), r.stats["nitref"], \
        r.stats["nrstrt"]


def start(n, seed=7):
    return np.random.default_rng(seed).uniform(-1, 1, n)


@pytest.mark.parametrize("which", ["LA", "SA", "LM", "SM", "BE"])
def test_sym_matches_oracle(which):
    nx, ny, nev, ncv = 17, 13, 5, 18
    A = laplace2d(nx, ny)
    n = nx * ny
    r0 = start(n)
    a = HostDouble().solve(lambda x: A @ x, n, nev, ncv, which, tol=1e-10, mxiter=3000, resid=r0)
    b = Oracle().solve(lambda x: A @ x, n, nev, ncv, which, tol=1e-10, mxiter=3000, resid=r0)
    assert a.info == b.info == 0 and a.ierr == b.ierr == 0
    assert counts(a) == counts(b)
    assert np.abs(a.d - b.d).max() / np.abs(b.d).max() < 1e-12
    # Ritz vectors agree up to sign
    for i in range(nev):
        assert min(np.linalg.norm(a.z[i] - b.z[i]), np.linalg.norm(a.z[i] + b.z[i])) < 1e-7
    # workl after *eupd: Ritz values (ipntr(8)) and error bounds (ipntr(9)) as dseupd.f:356-412 documents
    ihd, ihb = a.ipntr_eupd[7] - 1, a.ipntr_eupd[8] - 1
    assert np.allclose(a.workl_eupd[ihd:ihd + nev], b.workl_eupd[ihd:ihd + nev], rtol=1e-12)
    assert np.allclose(a.workl_eupd[ihb:ihb + nev], b.workl_eupd[ihb:ihb + nev], rtol=1e-3, atol=1e-18)


def test_sym_fixed_restart_budget_state_matches():
    """with a fixed restart budget (expect info=1) H, Ritz values and bounds must agree."""
    nx, ny, nev, ncv = 40, 33, 6, 20
    A = laplace2d(nx, ny)
    n = nx * ny
    r0 = start(n, 3)
    a = HostDouble().solve(lambda x: A @ x, n, nev, ncv, "LA", tol=1e-12, mxiter=3, resid=r0, eupd=False)
    b = Oracle().solve(lambda x: A @ x, n, nev, ncv, "LA", tol=1e-12, mxiter=3, resid=r0, eupd=False)
    assert a.info == b.info == 1
    assert counts(a) == counts(b)
    ih, ritz, bnd = a.ipntr[4] - 1, a.ipntr[5] - 1, a.ipntr[6] - 1
    assert np.allclose(a.workl[ih:ih + 2 * ncv], b.workl[ih:ih + 2 * ncv], rtol=1e-9, atol=1e-11)
    assert np.allclose(a.workl[ritz:ritz + ncv], b.workl[ritz:ritz + ncv], rtol=1e-11)
    assert np.allclose(a.workl[bnd:bnd + ncv], b.workl[bnd:bnd + ncv], rtol=1e-5, atol=1e-14)


def test_dssimp_eigenvalues():
    nx = 10
    r = HostDouble().solve(dssimp_av(nx), nx * nx, 4, 20, "LM", tol=0.0, mxiter=300)
    assert r.info == 0 and r.ierr == 0 and r.nconv == 4
    assert np.abs(r.d - dssimp_exact(nx, 4)).max() < 1e-10


@pytest.mark.parametrize("which", ["LM", "SM", "LR", "SR"])  # real simple spectrum: LI/SI would be all ties
def test_nonsym_matches_oracle(which):
    nx, nev, ncv = 12, 4, 16
    A = convdiff2d(nx, 7.0)
    n = nx * nx
    r0 = start(n, 11)
    a = HostDouble().solve(lambda x: A @ x, n, nev, ncv, which, sym=False, tol=1e-10, mxiter=3000, resid=r0)
    b = Oracle().solve(lambda x: A @ x, n, nev, ncv, which, sym=False, tol=1e-10, mxiter=3000, resid=r0)
    assert a.info == b.info and a.ierr == b.ierr == 0
    assert counts(a) == counts(b)
    la, lb = a.dr[:a.nconv] + 1j * a.di[:a.nconv], b.dr[:b.nconv] + 1j * b.di[:b.nconv]
    assert np.abs(la - lb).max() / np.abs(lb).max() < 1e-10
    for i in range(a.nconv):
        assert min(np.linalg.norm(a.z[i] - b.z[i]), np.linalg.norm(a.z[i] + b.z[i])) < 1e-6


@pytest.mark.parametrize("which", ["LM", "LI", "SR"])
def test_nonsym_complex_pairs(which):
    """Complex conjugate Ritz pairs: double-shift sweeps (dnapps.f:455-530), pair-preserving selection
    (dngets.f:191-195), complex eigenvector handling in dneupd."""
    rng = np.random.default_rng(5)
    n = 120
    A = rng.standard_normal((n, n)) / np.sqrt(n) + np.diag(np.linspace(0, 3, n))
    r0 = start(n, 2)
    a = HostDouble().solve(lambda x: A @ x, n, 5, 24, which, sym=False, tol=1e-10, mxiter=3000, resid=r0)
    b = Oracle().solve(lambda x: A @ x, n, 5, 24, which, sym=False, tol=1e-10, mxiter=3000, resid=r0)
    assert a.info == b.info == 0 and a.ierr == b.ierr == 0
    assert counts(a) == counts(b)
    la, lb = a.dr[:a.nconv] + 1j * a.di[:a.nconv], b.dr[:b.nconv] + 1j * b.di[:b.nconv]
    assert np.abs(la - lb).max() < 1e-9
    ev = np.linalg.eigvals(A)
    k = 0
    while k < a.nconv:
        assert np.abs(ev - la[k]).min() < 1e-8
        if a.di[k] != 0:
            x = a.z[k] + 1j * a.z[k + 1]
            assert np.linalg.norm(A @ x - la[k] * x) < 1e-8
            k += 2
        else:
            assert np.linalg.norm(A @ a.z[k] - a.dr[k] * a.z[k]) < 1e-8
            k += 1


def test_float32_matches_oracle():
    nx, ny, nev, ncv = 15, 12, 4, 14
    A = laplace2d(nx, ny).astype(np.float32)
    n = nx * ny
    r0 = start(n).astype(np.float32)
    a = HostDouble().solve(lambda x: A @ x, n, nev, ncv, "LA", tol=1e-5, mxiter=3000, resid=r0, dtype=np.float32)
    b = Oracle().solve(lambda x: A @ x, n, nev, ncv, "LA", tol=1e-5, mxiter=3000, resid=r0, dtype=np.float32)
    assert a.info == b.info == 0 and a.nconv == b.nconv
    assert np.abs(a.d - b.d).max() / np.abs(b.d).max() < 1e-4


def test_generalized_and_shift_invert_modes():
    """(f)-row functionality: bmat='G' hand-offs (ido=2), modes 2 and 3, purification (dseupd.f:840-857)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla
    n = 100
    A = sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]).tocsc()
    M = sp.diags([np.ones(n - 1) / 6, 4 * np.ones(n) / 6, np.ones(n - 1) / 6], [-1, 0, 1]).tocsc()
    lu = sla.splu(A)
    lu2 = sla.splu(A.tocsc())
    luM = sla.splu(M)
    r0 = start(n, 4)
    cases = [
        dict(op=lambda x: lu.solve(x), mode=3, bmat="I", sigma=0.0, bop=None),
        dict(op=lambda x, is_bx=False: lu2.solve(x if is_bx else M @ x), mode=3, bmat="G", sigma=0.0,
             bop=lambda x: M @ x),
        dict(op=lambda x: (luM.solve(A @ x), A @ x), mode=2, bmat="G", sigma=0.0, bop=lambda x: M @ x),
    ]
    for c in cases:
        a = HostDouble().solve(c["op"], n, 4, 12, "LM", tol=1e-12, mxiter=500, mode=c["mode"], bmat=c["bmat"],
                               sigma=c["sigma"], bop=c["bop"], resid=r0)
        b = Oracle().solve(c["op"], n, 4, 12, "LM", tol=1e-12, mxiter=500, mode=c["mode"], bmat=c["bmat"],
                           sigma=c["sigma"], bop=c["bop"], resid=r0)
        assert a.info == b.info == 0 and a.ierr == b.ierr == 0, c
        assert counts(a) == counts(b), c
        assert np.abs(a.d - b.d).max() / np.abs(b.d).max() < 1e-11
        for i in range(4):
            assert min(np.linalg.norm(a.z[i] - b.z[i]), np.linalg.norm(a.z[i] + b.z[i])) < 1e-7


def test_user_supplied_shifts_ido3():
    """ishift = 0: the RCI returns ido=3 and reads np shifts from workl(ipntr(11)) (dsaup2.f:713-743)."""
    import ctypes as C
    nx, ny, nev, ncv = 12, 9, 3, 12
    A = laplace2d(nx, ny)
    n = nx * ny

    def run(cls):
        o = cls()
        L, p = o.L, "d"
        v = np.zeros((ncv, n)); workd = np.zeros(3 * n); workl = np.zeros(ncv * ncv + 8 * ncv)
        iparam = np.zeros(11, dtype=np.int32); ipntr = np.zeros(14, dtype=np.int32)
        iparam[[0, 2, 3, 6]] = [0, 200, 1, 1]
        ido = C.c_int(0); info = C.c_int(1); tol = C.c_double(1e-10)
        resid = start(n, 9)
        aupd = o._fn("dsaupd")
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        n3 = 0
        while True:
            aupd(*o._ctxargs(p), C.byref(ido), b"I", n, b"LA", nev, C.byref(tol), dp(resid), ncv, dp(v), n, ip(iparam),
                 ip(ipntr), dp(workd), dp(workl), len(workl), C.byref(info))
            if ido.value in (-1, 1):
                workd[ipntr[1] - 1:ipntr[1] - 1 + n] = A @ workd[ipntr[0] - 1:ipntr[0] - 1 + n]
            elif ido.value == 3:
                npsh = iparam[7]
                ritz = workl[ipntr[5] - 1:ipntr[5] - 1 + ncv]
                # exact shifts supplied by hand: the np unwanted Ritz values (smallest for 'LA')
                workl[ipntr[10] - 1:ipntr[10] - 1 + npsh] = np.sort(ritz)[:npsh]
                n3 += 1
            else:
                break
        return info.value, iparam.copy(), workl[ipntr[5] - 1:ipntr[5] - 1 + nev].copy(), n3
    a, b = run(HostDouble), run(Oracle)
    assert a[0] == b[0] == 0 and a[3] == b[3] > 0
    assert np.array_equal(a[1], b[1])
    assert np.allclose(a[2], b[2], rtol=1e-12)
    ev = np.sort(np.linalg.eigvalsh(A.toarray()))
    assert np.abs(np.sort(a[2]) - ev[-nev:]).max() < 1e-8


@pytest.mark.parametrize("kw,expect", [
    (dict(n=0), -1), (dict(nev=0), -2), (dict(ncv=3, nev=3), -3), (dict(ncv=200), -3), (dict(mxiter=0), -4),
    (dict(which="XX"), -5), (dict(bmat="X"), -6), (dict(lworkl_short=True), -7), (dict(mode=6), -10),
    (dict(mode=1, bmat="G"), -11), (dict(ishift=2), -12), (dict(nev=1, which="BE", ncv=5), -13)])
def test_sym_argument_errors_match_reference(kw, expect):
    """dsaupd.f:501-543: argument errors come back as info<0 with ido=99, never an abort."""
    import ctypes as C
    base = dict(n=100, nev=3, ncv=10, which="LM", bmat="I", mode=1, ishift=1, mxiter=10)
    base.update({k: v for k, v in kw.items() if k != "lworkl_short"})
    for cls in (HostDouble, Oracle):
        o = cls()
        n, nev, ncv = base["n"], base["nev"], base["ncv"]
        nn = max(n, 1)
        lworkl = ncv * ncv + 8 * ncv - (1 if kw.get("lworkl_short") else 0)
        v = np.zeros(nn * ncv); workd = np.zeros(3 * nn); workl = np.zeros(ncv * ncv + 8 * ncv)
        resid = np.zeros(nn)
        iparam = np.zeros(11, dtype=np.int32); ipntr = np.zeros(14, dtype=np.int32)
        iparam[[0, 2, 3, 6]] = [base["ishift"], base["mxiter"], 1, base["mode"]]
        ido = C.c_int(0); info = C.c_int(0); tol = C.c_double(0.0)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        o._fn("dsaupd")(*o._ctxargs("d"), C.byref(ido), base["bmat"].encode(), n, base["which"].encode(), nev,
                        C.byref(tol), dp(resid), ncv, dp(v), nn, ip(iparam), ip(ipntr), dp(workd), dp(workl), lworkl,
                        C.byref(info))
        assert ido.value == 99 and info.value == expect, (cls.__name__, info.value)


@pytest.mark.parametrize("kw,expect", [
    (dict(ncv=4, nev=3), -3), (dict(which="LA"), -5), (dict(mode=5), -10), (dict(lworkl_short=True), -7)])
def test_nonsym_argument_errors_match_reference(kw, expect):
    import ctypes as C
    base = dict(n=100, nev=3, ncv=10, which="LM", bmat="I", mode=1, ishift=1, mxiter=10)
    base.update({k: v for k, v in kw.items() if k != "lworkl_short"})
    for cls in (HostDouble, Oracle):
        o = cls()
        n, nev, ncv = base["n"], base["nev"], base["ncv"]
        lworkl = 3 * ncv * ncv + 6 * ncv - (1 if kw.get("lworkl_short") else 0)
        v = np.zeros(n * ncv); workd = np.zeros(3 * n); workl = np.zeros(3 * ncv * ncv + 6 * ncv); resid = np.zeros(n)
        iparam = np.zeros(11, dtype=np.int32); ipntr = np.zeros(14, dtype=np.int32)
        iparam[[0, 2, 3, 6]] = [base["ishift"], base["mxiter"], 1, base["mode"]]
        ido = C.c_int(0); info = C.c_int(0); tol = C.c_double(0.0)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        o._fn("dnaupd")(*o._ctxargs("d"), C.byref(ido), base["bmat"].encode(), n, base["which"].encode(), nev,
                        C.byref(tol), dp(resid), ncv, dp(v), n, ip(iparam), ip(ipntr), dp(workd), dp(workl), lworkl,
                        C.byref(info))
        assert ido.value == 99 and info.value == expect, (cls.__name__, info.value)


def test_invariant_subspace_restart():
    """Start vector inside a 3-dimensional invariant subspace: the factorisation breaks down (rnorm = 0) and the
    reference regenerates a vector with dgetv0 (dsaitr.f:378-427, nrstrt > 0)."""
    n = 60
    diag = np.arange(1, n + 1, dtype=float)
    r0 = np.zeros(n)
    r0[[3, 17, 41]] = [1.0, -2.0, 0.5]
    a = HostDouble().solve(lambda x: diag * x, n, 4, 12, "LM", tol=1e-10, mxiter=500, resid=r0)
    b = Oracle().solve(lambda x: diag * x, n, 4, 12, "LM", tol=1e-10, mxiter=500, resid=r0)
    assert b.stats["nrstrt"] > 0
    assert a.info == b.info and a.stats["nrstrt"] == b.stats["nrstrt"]
    assert a.nconv == b.nconv
    assert np.abs(np.sort(a.d) - np.sort(b.d)).max() < 1e-8


def test_max_iterations_info_1():
    nx, ny = 30, 29
    A = laplace2d(nx, ny)
    r0 = start(nx * ny)
    for cls in (HostDouble, Oracle):
        r = cls().solve(lambda x: A @ x, nx * ny, 6, 14, "SA", tol=1e-14, mxiter=2, resid=r0, eupd=False)
        assert r.info == 1 and int(r.iparam[2]) == 3


@pytest.mark.parametrize("nranks", [2, 3])
def test_parpack_semantics_match_oracle(nranks):
    """pdsaupd/pdseupd path of the host logic (per-rank seeds, no initial OP*x, fused all-reduces) against the
    oracle's PARPACK mode, both on P logical ranks; the known answer is icb_parpack_c.c's 992..1000."""
    N = 1000
    cnts, offs = split_rows(N, nranks)

    def run(cls):
        world = LogicalRanks(nranks)

        def rank_main(r, ar):
            diag = np.arange(offs[r] + 1, offs[r + 1] + 1, dtype=float)
            return cls(rank=r, nranks=nranks, allreduce=ar).solve(lambda x: diag * x, cnts[r], 9, 19, "LM", tol=1e-6,
                                                                   mxiter=10000, c_abi_tol=True)
        return world.run(rank_main)
    a, b = run(HostDouble), run(Oracle)
    for ra, rb in zip(a, b):
        assert ra.info == rb.info == 0 and ra.ierr == rb.ierr == 0
        assert np.abs(ra.d - np.arange(992, 1001)).max() < 1e-5
        assert counts(ra) == counts(rb)
        assert np.abs(ra.d - rb.d).max() < 1e-9


def test_parpack_nonsym_semantics_match_oracle():
    nranks = 2
    nx = 14
    A = convdiff2d(nx, 5.0).tocsr()
    n = nx * nx
    cnts, offs = split_rows(n, nranks)
    r0 = start(n, 21)

    def run(cls):
        world = LogicalRanks(nranks)
        xs = [None] * nranks

        def rank_main(r, ar):
            def op(x):
                # "halo exchange": gather the full vector through the all-reduce callback
                full = np.zeros(n)
                full[offs[r]:offs[r + 1]] = x
                full = ar(full, 0)
                return (A @ full)[offs[r]:offs[r + 1]]
            return cls(rank=r, nranks=nranks, allreduce=ar).solve(op, cnts[r], 4, 16, "LM", sym=False, tol=1e-10,
                                                                   mxiter=2000, resid=r0[offs[r]:offs[r + 1]])
        return world.run(rank_main)
    a, b = run(HostDouble), run(Oracle)
    for ra, rb in zip(a, b):
        assert ra.info == rb.info == 0
        assert counts(ra) == counts(rb)
        assert np.abs(ra.dr - rb.dr).max() / np.abs(rb.dr).max() < 1e-10


@pytest.mark.parametrize("sym,fused", [(True, False), (True, True), (False, False), (False, True)])
def test_registered_operator_runs_whole_solve_in_one_call(sym, fused):
    """§8(f) row 2: with a registered OP the solve never hands off (no ido=+-1) and gives the RCI result bit for bit
    (the control flow is the same code; only the yield is replaced by the operator call)."""
    nx = 14
    if sym:
        A, nev, ncv, which = laplace2d(nx, nx + 3), 4, 16, "LA"
    else:
        A, nev, ncv, which = convdiff2d(nx, rho=10.0), 4, 16, "LM"
    n = A.shape[0]
    r0 = start(n, 11)
    op = lambda x: A @ x
    a = HostDouble().solve(op, n, nev, ncv, which, sym=sym, tol=1e-10, mxiter=500, resid=r0)
    hd = HostDouble()
    hd.register_op(op, n, fused=fused)
    b = hd.solve(None, n, nev, ncv, which, sym=sym, tol=1e-10, mxiter=500, resid=r0)
    assert a.info == b.info == 0 and a.ierr == b.ierr == 0
    assert b.nsteps == 0 and a.nsteps > 0            # one *aupd call, no hand-off
    assert a.stats == b.stats and counts(a) == counts(b)
    assert np.array_equal(a.v, b.v) and np.array_equal(a.workl, b.workl)
    if sym:
        assert np.array_equal(a.d, b.d)
    else:
        assert np.array_equal(a.dr, b.dr) and np.array_equal(a.di, b.di)
    if fused:
        assert 0.0 <= hd.fused_dot_maxdiff(sym) < 1e-12


def test_registered_operator_is_ignored_outside_mode1():
    """The registration applies to mode 1 / bmat='I' only; shift-invert keeps the RCI hand-offs."""
    import scipy.sparse.linalg as spla
    nx = 12
    A = laplace2d(nx, nx)
    n = A.shape[0]
    lu = spla.splu(A.tocsc())
    hd = HostDouble()
    hd.register_op(lambda x: A @ x, n, fused=True)
    r = hd.solve(lambda x: lu.solve(x), n, 4, 12, "LM", tol=1e-10, mxiter=300, mode=3, sigma=0.0, resid=start(n))
    assert r.info == 0 and r.nsteps > 0
    assert np.allclose(np.sort(r.d), np.sort(np.linalg.eigvalsh(A.toarray()))[:4], rtol=1e-9)


@pytest.mark.parametrize("sym", [True, False])
@pytest.mark.parametrize("registered", [False, True])
def test_device_resident_sweeps_change_nothing(sym, registered):
    """IrlBase::extend, deferred mode: a whole sweep of steps is enqueued without a host round trip (each step has its
    own mailbox slot, the next step's scale and the rare-path tests are formed from the slot by the gated start of
    step) and the host replays its bookkeeping once per sweep.  Path and results must be those of the synchronous
    mode and of the oracle, with one round trip per sweep instead of one per step."""
    if sym:
        A, nev, ncv, which = laplace2d(19, 16), 4, 16, "LA"
    else:
        A, nev, ncv, which = convdiff2d(15, rho=10.0), 4, 16, "LM"
    n = A.shape[0]
    r0 = start(n, 5)
    hd, hs = HostDouble(), HostDouble()
    hs.set_deferral(False)
    if registered:
        hd.register_op(lambda x: A @ x, n, fused=True)
        hs.register_op(lambda x: A @ x, n, fused=True)
    a = hd.solve(lambda x: A @ x, n, nev, ncv, which, sym=sym, tol=1e-10, mxiter=500, resid=r0)
    s = hs.solve(lambda x: A @ x, n, nev, ncv, which, sym=sym, tol=1e-10, mxiter=500, resid=r0)
    b = Oracle().solve(lambda x: A @ x, n, nev, ncv, which, sym=sym, tol=1e-10, mxiter=500, resid=r0)
    assert a.info == b.info == s.info == 0 and counts(a) == counts(b) == counts(s)
    assert np.array_equal(a.workl, s.workl)           # bit-identical projected problem
    steps, trips, trips_rt = hd.deferred_stats(sym)
    nopx, sweeps = int(a.iparam[8]), int(a.iparam[2]) + 1
    assert trips == 0
    assert steps >= nopx - 1                          # everything but the start-vector product ran deferred
    assert trips_rt <= 3 * sweeps + 4                 # one block read per sweep (+ the restart norm, + getv0)
    assert hs.deferred_stats(sym)[0] == 0 and hs.deferred_stats(sym)[2] >= nopx - 1


def test_device_resident_sweep_is_cut_short_by_a_breakdown():
    """Start vector inside a 3-dimensional invariant subspace: rnorm collapses in the middle of a sweep.  The gated
    start of the next step trips, every later kernel of the batch exits at once, and the host resumes with the
    reference's restart logic (dsaitr.f:378-427) from exactly that step -- same path as the oracle and as the
    synchronous mode."""
    n = 60
    diag = np.arange(1, n + 1, dtype=float)
    r0 = np.zeros(n)
    r0[[3, 17, 41]] = [1.0, -2.0, 0.5]
    hd, hs = HostDouble(), HostDouble()
    hs.set_deferral(False)
    a = hd.solve(lambda x: diag * x, n, 4, 12, "LM", tol=1e-10, mxiter=500, resid=r0)
    s = hs.solve(lambda x: diag * x, n, 4, 12, "LM", tol=1e-10, mxiter=500, resid=r0)
    b = Oracle().solve(lambda x: diag * x, n, 4, 12, "LM", tol=1e-10, mxiter=500, resid=r0)
    assert b.stats["nrstrt"] > 0
    assert counts(a) == counts(s)
    assert a.info == b.info and a.stats["nrstrt"] == b.stats["nrstrt"] and a.nconv == b.nconv
    assert np.array_equal(a.workl, s.workl)
    assert hd.deferred_stats()[1] >= 1                # at least one batch was cut short
    assert np.abs(np.sort(a.d) - np.sort(b.d)).max() < 1e-8


def _buckling_cayley_problem(n=80):
    import scipy.sparse as sp
    K = sp.diags([-np.ones(n - 1), 2.2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]).tocsc()            # SPD
    KG = sp.diags([0.3 * np.ones(n - 1), np.linspace(-1.0, 2.0, n), 0.3 * np.ones(n - 1)], [-1, 0, 1]).tocsc()  # indefinite
    M = sp.diags([np.ones(n - 1) / 6, 4 * np.ones(n) / 6, np.ones(n - 1) / 6], [-1, 0, 1]).tocsc()    # SPD
    return K, KG, M


@pytest.mark.parametrize("backend", ["oracle", "hostdouble"])
def test_buckling_and_cayley_modes(backend):
    """dsaupd modes 4 and 5 (dsaupd.f:118-140) and their back-transforms in dseupd (dseupd.f:672-712, 796-838):
    lambda = sigma*theta/(theta-1) (buckling), lambda = sigma*(theta+1)/(theta-1) (Cayley), against dense solutions;
    the product's host logic must follow the oracle count for count."""
    import scipy.linalg as sl
    import scipy.sparse.linalg as sla
    n = 80
    K, KG, M = _buckling_cayley_problem(n)
    r0 = start(n, 6)
    cls = Oracle if backend == "oracle" else HostDouble
    # ---- mode 4: K x = lambda KG x, OP = inv(K - sigma KG) K, B = K ----
    sigma = 0.6
    lu = sla.splu((K - sigma * KG).tocsc())
    lam = sl.eig(K.toarray(), KG.toarray(), right=False)
    lam = np.sort(lam[np.isfinite(lam)].real)

    def op4(x, is_bx=False):
        return lu.solve(x if is_bx else K @ x)
    r = cls().solve(op4, n, 4, 16, "LM", tol=1e-12, mxiter=500, mode=4, bmat="G", sigma=sigma, bop=lambda x: K @ x,
                    resid=r0)
    assert r.info == 0 and r.ierr == 0 and r.nconv == 4
    near = lam[np.argsort(-np.abs(lam / (lam - sigma)))[:4]]          # 'LM' in theta = lambda/(lambda-sigma)
    assert np.abs(np.sort(r.d) - np.sort(near)).max() < 1e-8
    for k in range(4):  # K z = lambda KG z
        assert np.linalg.norm(K @ r.z[k] - r.d[k] * (KG @ r.z[k])) < 1e-7 * np.linalg.norm(K @ r.z[k])
    ref4 = r
    # ---- mode 5: A x = lambda M x, OP = inv(A - sigma M)(A + sigma M), B = M ----
    A = K
    sigma5 = 1.5
    lu5 = sla.splu((A - sigma5 * M).tocsc())
    gev = np.sort(sl.eigh(A.toarray(), M.toarray(), eigvals_only=True))

    def op5(x, bx):
        return lu5.solve(A @ x + sigma5 * (bx if bx is not None else M @ x))
    r = cls().solve(op5, n, 4, 16, "LM", tol=1e-12, mxiter=500, mode=5, bmat="G", sigma=sigma5, bop=lambda x: M @ x,
                    resid=r0)
    assert r.info == 0 and r.ierr == 0 and r.nconv == 4
    near = gev[np.argsort(-np.abs((gev + sigma5) / (gev - sigma5)))[:4]]   # 'LM' in the Cayley theta
    assert np.abs(np.sort(r.d) - np.sort(near)).max() < 1e-8
    for k in range(4):
        assert np.linalg.norm(A @ r.z[k] - r.d[k] * (M @ r.z[k])) < 1e-7 * np.linalg.norm(A @ r.z[k])
    if backend == "hostdouble":
        o4 = Oracle().solve(op4, n, 4, 16, "LM", tol=1e-12, mxiter=500, mode=4, bmat="G", sigma=sigma,
                            bop=lambda x: K @ x, resid=r0)
        o5 = Oracle().solve(op5, n, 4, 16, "LM", tol=1e-12, mxiter=500, mode=5, bmat="G", sigma=sigma5,
                            bop=lambda x: M @ x, resid=r0)
        assert counts(ref4) == counts(o4) and counts(r) == counts(o5)
        assert np.abs(ref4.d - o4.d).max() < 1e-10 and np.abs(r.d - o5.d).max() < 1e-10


import golden_cases  # noqa: E402


@pytest.mark.parametrize("c", golden_cases.load(), ids=golden_cases.case_id)
def test_host_logic_reproduces_committed_scipy_arpack_vectors(c):
    """The product's control code (over the test double) against golden vectors made by a third implementation."""
    A = golden_cases.PROBLEMS[c["problem"]]()
    n = A.shape[0]
    r = HostDouble().solve(lambda x: A @ x, n, c["nev"], c["ncv"], c["which"], sym=c["sym"], tol=c["tol"], mxiter=3000,
                           resid=golden_cases.start_vector(c, n))
    golden_cases.check_against_golden(c, r, c["nev"])


@pytest.mark.parametrize("mode", [3, 4])
def test_nonsym_complex_shift_real_and_imaginary_part_modes(mode):
    """dnaupd modes 3 and 4 with a complex shift (dnaupd.f:119-143, EXAMPLES/NONSYM/dndrv4.f style): OP is the real
    (mode 3) or imaginary (mode 4) part of inv(A - sigma I); dneupd leaves the Ritz values of OP untransformed
    (type REALPT / IMAGPT, dneupd.f:950-991) and the caller recovers lambda by Rayleigh quotients (remark 3)."""
    n = 60
    A = np.diag(2.0 + 0.05 * np.arange(n)) + np.diag(-1.3 * np.ones(n - 1), -1) + np.diag(0.7 * np.ones(n - 1), 1) \
        + np.diag(0.4 * np.ones(n - 3), 3)
    sigma = 2.5 + 0.8j
    S = np.linalg.inv(A - sigma * np.eye(n))
    op = (lambda x: (S @ x).real) if mode == 3 else (lambda x: (S @ x).imag)
    r0 = start(n, 9)
    kw = dict(sym=False, tol=1e-9, mxiter=3000, mode=mode, bmat="I", sigma=sigma.real, sigmai=sigma.imag, resid=r0)
    a = HostDouble().solve(op, n, 4, 24, "LM", **kw)
    b = Oracle().solve(op, n, 4, 24, "LM", **kw)
    assert a.info == b.info == 0 and a.ierr == b.ierr == 0
    assert counts(a) == counts(b)
    assert np.abs(a.dr[:a.nconv] - b.dr[:b.nconv]).max() < 1e-8 and np.abs(a.di[:a.nconv] - b.di[:b.nconv]).max() < 1e-8
    # Rayleigh quotients of the returned vectors are eigenvalues of A (the ones that dominate OP's spectrum)
    ev = np.linalg.eigvals(A)
    k = 0
    while k < a.nconv:
        if a.di[k] != 0 and k + 1 < a.nconv + 1:
            x = a.z[k] + 1j * a.z[k + 1]
            k += 2
        else:
            x = a.z[k].astype(complex)
            k += 1
        lam = (x.conj() @ (A @ x)) / (x.conj() @ x)
        assert np.abs(ev - lam).min() < 1e-6 or np.abs(ev - np.conj(lam)).min() < 1e-6


@pytest.mark.parametrize("scale", [1e200, 1e-200])
def test_start_vector_whose_squares_overflow_or_underflow(scale):
    """pdnorm2.f:72-80 and the BLAS dnrm2 behind dgetv0 divide by the largest entry before squaring; this library's
    norms are plain sums of squares, so a start vector of entries ~1e200 (or ~1e-200) is first scaled by a power of two
    (IrlBase::rescale_start_vector).  The first Lanczos vector is the same either way: same path, same eigenvalues as the
    oracle -- and as the run from the unscaled vector."""
    A = laplace2d(15, 14)
    n = A.shape[0]
    r0 = start(n, 21)
    base = HostDouble().solve(lambda x: A @ x, n, 4, 14, "LA", tol=1e-10, mxiter=500, resid=r0)
    a = HostDouble().solve(lambda x: A @ x, n, 4, 14, "LA", tol=1e-10, mxiter=500, resid=r0 * scale)
    b = Oracle().solve(lambda x: A @ x, n, 4, 14, "LA", tol=1e-10, mxiter=500, resid=r0 * scale)
    assert a.info == b.info == base.info == 0
    assert counts(a) == counts(b)
    assert np.abs(a.d - b.d).max() <= 1e-12 * np.abs(b.d).max()
    assert np.abs(a.d - base.d).max() <= 1e-10 * np.abs(b.d).max()
