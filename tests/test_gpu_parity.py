"""GPU parity tests: the CUDA path (through the C-ABI of libarpack_b200.so) against the CPU oracle on the
same seeded inputs.  Floating-point tolerances are north_star's: eigenvalues 1e-10 relative (FP64) / 1e-4 (FP32),
||A x - lambda x|| <= tol*||A||; convergence counts (nconv, restarts, OP*x, re-orth steps) identical."""
import ctypes as C

import numpy as np
import pytest

from backends import Oracle, lib as oracle_lib
from problems import convdiff2d, dssimp_av, dssimp_exact, laplace2d, laplace3d

pytestmark = pytest.mark.gpu

RTOL64 = 1e-10  # north_star: converged eigenvalues agree to 1e-10 relative in FP64
RTOL32 = 1e-4   # ... 1e-4 in FP32


@pytest.fixture(scope="module")
def ab():
    import arpack_ng_b200 as m
    m.lib()
    return m


def _torch():
    import torch
    return torch


def _counts(r):
    return int(r.nconv), int(r.iparam[2]), int(r.iparam[8]), int(r.iparam[10])


# --------------------------------------------------------------------------------------------------
# known-answer tests of the reference, run through the C-ABI
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("host_buffers", [False, True])
def test_icb_arpack_c_diag(ab, host_buffers):
    """TESTS/icb_arpack_c.c:31-91: A = diag(1..1000), nev=9, ncv=19, 'LM', tol=1e-6 -> d = 992..1000 (1e-5)."""
    torch = _torch()
    n = 1000
    if host_buffers:
        diag = np.arange(1, n + 1, dtype=np.float64)

        def op(x, y, *_):
            y[:] = diag * x
    else:
        diag = torch.arange(1, n + 1, dtype=torch.float64, device="cuda")

        def op(x, y, *_):
            torch.mul(diag, x, out=y)
    ab.lib().ab200_reset_seed()  # the dgetv0 seed is SAVE'd across solves (dgetv0.f:164): start a fresh "process"
    r = ab.solve(op, n, 9, 19, "LM", tol=1e-6, mxiter=10000, host_buffers=host_buffers)
    assert r.info == 0 and r.ierr == 0
    assert r.nconv >= 9
    assert np.abs(r.d - np.arange(992, 1001)).max() < 1e-5
    ref = Oracle().solve(lambda x: np.arange(1, n + 1) * x, n, 9, 19, "LM", tol=1e-6, mxiter=10000, c_abi_tol=True)
    assert _counts(r) == _counts(ref)


def test_bug_1315_double(ab):
    """TESTS/bug_1315_double.c:23-84: dnaupd_c/dneupd_c on the same matrix, tol=0 -> dr[i] = 1000-i (1e-6)."""
    torch = _torch()
    n = 1000
    diag = torch.arange(1, n + 1, dtype=torch.float64, device="cuda")
    r = ab.solve(lambda x, y, *_: torch.mul(diag, x, out=y), n, 9, 19, "LM", sym=False, tol=0.0, mxiter=10 * n)
    assert r.info == 0 and r.ierr == 0
    assert np.abs(r.dr[:9] - (1000 - np.arange(9))).max() < 1e-6
    assert np.abs(r.di[:9]).max() == 0.0


def test_bug_1315_single(ab):
    """TESTS/bug_1315_single.c:75: float twin, 1e-1."""
    torch = _torch()
    n = 1000
    diag = torch.arange(1, n + 1, dtype=torch.float32, device="cuda")
    r = ab.solve(lambda x, y, *_: torch.mul(diag, x, out=y), n, 9, 19, "LM", sym=False, tol=0.0, mxiter=10 * n,
                 dtype=np.float32)
    assert r.info == 0 and r.ierr == 0
    assert np.abs(r.dr[:9] - (1000 - np.arange(9))).max() < 1e-1


def test_dssimp_config1(ab):
    """BASELINE config 1 = EXAMPLES/SIMPLE/dssimp.f: nx=10, nev=4, ncv=20, 'LM', tol=0, mxiter=300, random start
    from the LAPACK dlarnv stream (info=0).  The spectrum has a double eigenvalue (919.78...), so restart counts are
    rounding-sensitive; the eigenvalues themselves are pinned by the analytic spectrum."""
    torch = _torch()
    nx = 10
    A = ab.CsrOperator.laplace2d(nx, nx, scale=float((nx + 1) ** 2))
    r = ab.solve(A, nx * nx, 4, 20, "LM", tol=0.0, mxiter=300)
    assert r.info == 0 and r.ierr == 0 and r.nconv == 4
    exact = dssimp_exact(nx, 4)
    assert np.abs(r.d - exact).max() / exact.max() < 1e-12
    rn = A.residuals(r.d, r.z, nx * nx)
    assert (rn / np.abs(r.d) < 1e-10).all()


def test_first_handoff_is_the_dlarnv_stream(ab):
    """dgetv0.f:236 draws the start vector from LAPACK dlarnv(idist=2, iseed={1,3,5,7}); the CUDA generator must
    reproduce that stream bit for bit (first hand-off: x = workd(ipntr(1)) = the random vector)."""
    torch = _torch()
    L = ab.lib()
    L.ab200_reset_seed()
    n, nev, ncv = 5000, 3, 12
    workl = np.zeros(ncv * ncv + 8 * ncv)
    iparam = np.zeros(11, dtype=np.int32)
    iparam[[0, 2, 3, 6]] = [1, 10, 1, 1]
    ipntr = np.zeros(14, dtype=np.int32)
    ido = np.zeros(1, dtype=np.int32)
    info = np.zeros(1, dtype=np.int32)
    resid = torch.zeros(n, dtype=torch.float64, device="cuda")
    v = torch.zeros(n * ncv, dtype=torch.float64, device="cuda")
    workd = torch.zeros(3 * n, dtype=torch.float64, device="cuda")
    ab.dsaupd_c(ido, "I", n, "LM", nev, 0.0, resid, ncv, v, n, iparam, ipntr, workd, workl, info)
    assert ido[0] == -1 and ipntr[0] == 1 and ipntr[1] == n + 1
    x = workd[:n].cpu().numpy()
    seed = np.array([1, 3, 5, 7], dtype=np.int32)
    ref = np.zeros(n)
    oracle_lib().ref_dlarnv2(seed.ctypes.data_as(C.POINTER(C.c_int)), n, ref.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.array_equal(x, ref)
    assert abs(ref[0] - 0.3957424639187579) < 1e-16  # SURVEY.md §8c golden value
    L.ab200_release(workl.ctypes.data)
    L.ab200_reset_seed()


# --------------------------------------------------------------------------------------------------
# oracle parity on seeded inputs (info = 1, hashed start vector)
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nx,ny,nev,ncv,which", [(48, 40, 6, 24, "LA"), (33, 29, 5, 20, "SA"), (64, 37, 8, 30, "LM"),
                                                   (25, 31, 4, 16, "BE")])
@pytest.mark.parametrize("mode", ["auto", "generic"])
def test_laplace2d_vs_oracle(ab, nx, ny, nev, ncv, which, mode):
    L = ab.lib()
    L.ab200_set_kernel_mode(0 if mode == "auto" else 1)
    try:
        A = ab.CsrOperator.laplace2d(nx, ny)
        n = A.n
        As = laplace2d(nx, ny)
        assert (abs(A.to_scipy() - As)).max() == 0.0
        r0 = ab.hashed_start_vector_numpy(n)
        assert np.array_equal(ab.hashed_start_vector(n).cpu().numpy(), r0)
        r = ab.solve(A, n, nev, ncv, which, tol=1e-10, mxiter=2000, resid=r0)
        ref = Oracle().solve(lambda x: As @ x, n, nev, ncv, which, tol=1e-10, mxiter=2000, resid=r0)
        assert r.info == ref.info == 0 and r.ierr == 0
        assert _counts(r) == _counts(ref)
        assert np.abs(r.d - ref.d).max() / np.abs(ref.d).max() < RTOL64
        rn = A.residuals(r.d, r.z, n)
        assert (rn <= 1e-10 * 8.0 * 10).all()
    finally:
        L.ab200_set_kernel_mode(0)


def test_laplace2d_host_buffers_match_device_buffers(ab):
    """The host-pointer path (unmodified CPU caller) and the device-pointer path run the same kernels."""
    nx, ny, nev, ncv = 40, 36, 5, 20
    A = ab.CsrOperator.laplace2d(nx, ny)
    r0 = ab.hashed_start_vector_numpy(A.n)
    rd = ab.solve(A, A.n, nev, ncv, "LA", tol=1e-10, mxiter=1000, resid=r0)
    rh = ab.solve(A, A.n, nev, ncv, "LA", tol=1e-10, mxiter=1000, resid=r0, host_buffers=True)
    assert _counts(rd) == _counts(rh)
    assert np.array_equal(rd.d, rh.d)
    assert np.allclose(rd.z[:A.n * nev].cpu().numpy(), rh.z[:A.n * nev], rtol=0, atol=1e-12)


def test_convdiff_nonsym_vs_oracle(ab):
    """BASELINE config 4 (dndrv1-style operator) at a size the oracle finishes: dnaupd nev=6 ncv=30 'LR'.
    rho*h/2 < 1 keeps the spectrum real and simple (as at the full size nx=2048, rho=100), so the wanted set is
    well defined and the counts must match the oracle exactly."""
    nx, nev, ncv, rho = 40, 6, 30, 10.0
    A = ab.CsrOperator.convdiff2d(nx, rho)
    As = convdiff2d(nx, rho)
    assert abs(A.to_scipy() - As).max() < 1e-9
    As = A.to_scipy()
    r0 = ab.hashed_start_vector_numpy(A.n)
    r = ab.solve(A, A.n, nev, ncv, "LR", sym=False, tol=1e-10, mxiter=2000, resid=r0)
    ref = Oracle().solve(lambda x: As @ x, A.n, nev, ncv, "LR", sym=False, tol=1e-10, mxiter=2000, resid=r0)
    assert r.info == ref.info == 0 and r.ierr == ref.ierr == 0
    assert _counts(r) == _counts(ref)
    lam, lam_ref = r.dr[:r.nconv] + 1j * r.di[:r.nconv], ref.dr[:ref.nconv] + 1j * ref.di[:ref.nconv]
    assert np.abs(np.sort_complex(lam) - np.sort_complex(lam_ref)).max() / np.abs(lam_ref).max() < RTOL64
    z = r.z[:A.n * r.nconv].cpu().numpy().reshape(r.nconv, A.n)
    for k in range(r.nconv):
        if r.di[k] == 0:
            assert np.linalg.norm(As @ z[k] - r.dr[k] * z[k]) < 1e-8 * abs(r.dr[k])


def test_convdiff_complex_spectrum(ab):
    """rho = 100 on a coarse grid (dnsimp.f:570): complex conjugate pairs with massively tied real parts, so the
    selected set is rounding dependent; every returned value must still be an eigenvalue of A and pairs must be
    conjugate (exercises the double-shift sweeps of dnapps.f:455-530 and the complex branch of dneupd)."""
    nx, nev, ncv = 12, 4, 20
    A = ab.CsrOperator.convdiff2d(nx, 100.0)
    As = A.to_scipy()
    ev = np.linalg.eigvals(As.toarray())
    r0 = ab.hashed_start_vector_numpy(A.n)
    r = ab.solve(A, A.n, nev, ncv, "SM", sym=False, tol=1e-10, mxiter=2000, resid=r0)
    assert r.info == 0 and r.ierr == 0 and r.nconv >= nev
    lam = r.dr[:r.nconv] + 1j * r.di[:r.nconv]
    assert np.abs(lam.imag).max() > 0
    for l in lam:
        assert np.abs(ev - l).min() < 1e-8 * abs(l)
    z = r.z[:A.n * (r.nconv + 1)].cpu().numpy().reshape(r.nconv + 1, A.n)
    k = 0
    while k < r.nconv:
        if r.di[k] != 0:  # (z_k, z_k+1) = (re, im) of the eigenvector of dr + i di
            x = z[k] + 1j * z[k + 1]
            assert np.linalg.norm(As @ x - lam[k] * x) < 1e-8 * abs(lam[k])
            assert r.di[k + 1] == -r.di[k]
            k += 2
        else:
            assert np.linalg.norm(As @ z[k] - r.dr[k] * z[k]) < 1e-8 * abs(r.dr[k])
            k += 1


def test_float32_sym_vs_oracle(ab):
    nx, ny, nev, ncv = 30, 26, 4, 16
    A64 = laplace2d(nx, ny)
    torch = _torch()
    A = ab.CsrOperator.from_scipy(A64)
    valf = A.val.float()
    n = A.n
    L = ab.lib()

    def op(x, y, *_):
        assert L.ab200_csr_spmv_f32(n, A.rowptr.data_ptr(), A.col.data_ptr(), valf.data_ptr(), x.data_ptr(),
                                    y.data_ptr()) == 0
    r0 = ab.hashed_start_vector_numpy(n).astype(np.float32)
    r = ab.solve(op, n, nev, ncv, "LA", tol=1e-5, mxiter=2000, resid=r0, dtype=np.float32)
    A32 = A64.astype(np.float32)
    ref = Oracle().solve(lambda x: A32 @ x, n, nev, ncv, "LA", tol=1e-5, mxiter=2000, resid=r0, dtype=np.float32)
    assert r.info == 0 and r.ierr == 0
    assert np.abs(r.d - ref.d).max() / np.abs(ref.d).max() < RTOL32
    assert r.nconv == ref.nconv


def test_laplace3d_generator_and_spmv(ab):
    torch = _torch()
    nx, ny, nz = 7, 5, 6
    A = ab.CsrOperator.laplace3d(nx, ny, nz)
    As = laplace3d(nx, ny, nz)
    assert abs(A.to_scipy() - As).max() == 0.0
    x = torch.randn(A.n, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    A(x, y)
    assert np.allclose(y.cpu().numpy(), As @ x.cpu().numpy(), rtol=1e-13, atol=1e-13)


def test_kernels_counted(ab):
    st = ab.launch_stats()
    assert st["kernels"] > 0


# --------------------------------------------------------------------------------------------------
# registered-operator mode: OP applied by the library, K1+K2+K3 fused
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["lap2d_small", "lap2d_tma", "lap3d", "convdiff", "lap2d_f32"])
def test_registered_csr_operator_matches_rci_and_oracle(ab, case):
    torch = _torch()
    sym, dtype, rtol = True, np.float64, RTOL64
    if case == "lap2d_small":
        A, nev, ncv, which = ab.CsrOperator.laplace2d(37, 29), 4, 16, "LA"
    elif case == "lap2d_tma":
        A, nev, ncv, which = ab.CsrOperator.laplace2d(300, 260), 6, 24, "LA"
    elif case == "lap3d":
        A, nev, ncv, which = ab.CsrOperator.laplace3d(24, 20, 22), 5, 20, "LA"
    elif case == "convdiff":
        A, nev, ncv, which, sym = ab.CsrOperator.convdiff2d(40, rho=10.0), 4, 20, "LM", False
    else:
        A, nev, ncv, which, dtype, rtol = ab.CsrOperator.laplace2d(64, 48), 4, 16, "LA", np.float32, RTOL32
        A = ab.CsrOperator(A.n, A.rowptr, A.col, A.val.float())
    n = A.n
    r0 = np.random.default_rng(5).uniform(-1, 1, n)
    tol = 1e-10 if dtype == np.float64 else 1e-5
    a = ab.solve(A, n, nev, ncv, which, sym=sym, tol=tol, mxiter=3000, resid=r0, dtype=dtype)
    b = ab.solve(None, n, nev, ncv, which, sym=sym, tol=tol, mxiter=3000, resid=r0, dtype=dtype, registered_op=A)
    assert a.info == b.info == 0 and a.ierr == b.ierr == 0
    assert b.nsteps == 0 and a.nsteps > 0          # the whole solve ran inside one *aupd_c call
    assert _counts(a) == _counts(b)
    if sym:
        assert np.abs(a.d - b.d).max() <= rtol * np.abs(a.d).max()
    else:
        assert np.abs(a.dr - b.dr).max() <= rtol * np.abs(a.dr).max()
        assert np.abs(a.di - b.di).max() <= rtol * max(np.abs(a.dr).max(), 1e-300)
    # the SpMV-epilogue dots (alpha, ||OP v||^2) agree with the CGS sweep of the same step
    assert b.fused_dot_maxdiff is not None and 0.0 <= b.fused_dot_maxdiff < (1e-11 if dtype == np.float64 else 1e-3)
    # and against the CPU oracle on the same start vector
    S = A.to_scipy().astype(dtype)
    ref = Oracle().solve(lambda x: S @ x, n, nev, ncv, which, sym=sym, tol=tol, mxiter=3000, resid=r0, dtype=dtype,
                         c_abi_tol=True)
    assert ref.info == 0
    if dtype == np.float64:
        assert _counts(b) == _counts(ref)
    if sym:
        assert np.abs(b.d - ref.d).max() <= rtol * np.abs(ref.d).max()
    else:
        assert np.abs(np.sort(b.dr[:nev]) - np.sort(ref.dr[:nev])).max() <= rtol * np.abs(ref.dr).max()


@pytest.mark.parametrize("pinned", [False, True])
def test_registered_host_csr_with_host_buffers_matches_device_path(ab, pinned):
    """A caller that owns EVERYTHING on the host (CSR matrix, resid, V, workd -- every caller of the reference) registers
    its host matrix: the library uploads it once, one dsaupd_c call runs the solve, V/resid come back at ido = 99 and
    dseupd_c works on the host arrays.  Same path (counts) and eigenvalues as the device-resident registered solve."""
    A = ab.CsrOperator.laplace2d(300, 260)
    n, nev, ncv = A.n, 6, 24
    r0 = np.random.default_rng(11).uniform(-1, 1, n)
    dev = ab.solve(None, n, nev, ncv, "LA", tol=1e-10, mxiter=3000, resid=r0, registered_op=A)
    H = ab.HostCsr.from_operator(A, pinned=pinned)
    hst = ab.solve(None, n, nev, ncv, "LA", tol=1e-10, mxiter=3000, resid=r0, registered_op=H, host_buffers=True,
                   pinned=pinned)
    assert dev.info == hst.info == 0 and dev.ierr == hst.ierr == 0
    assert hst.nsteps == 0                                  # never handed an OP*x back
    assert _counts(dev) == _counts(hst)
    assert np.abs(dev.d - hst.d).max() <= RTOL64 * np.abs(dev.d).max()
    # the eigenvectors arrived in the caller's HOST array: check A z = lambda z there
    S = A.to_scipy()
    Z = np.asarray(hst.z).reshape(ncv, n)[:nev].T
    res = np.linalg.norm(S @ Z - Z * hst.d[None, :], axis=0)
    assert (res <= 1e-8 * 8.0).all(), res


def test_registered_operator_row_count_mismatch_fails_loudly(ab):
    A = ab.CsrOperator.laplace2d(20, 20)
    with pytest.raises(ab.ArpackB200Error):   # info = -9990 from the C-ABI, surfaced by the binding
        ab.solve(None, A.n - 1, 3, 12, "LA", tol=1e-8, mxiter=10, registered_op=A, eupd=False)


def test_registration_is_one_shot(ab):
    """A registration is consumed by the solve it was made for: the next solve on the same workl address hands every
    OP*x back to the caller again."""
    A = ab.CsrOperator.laplace2d(31, 23)
    r0 = np.random.default_rng(2).uniform(-1, 1, A.n)
    for _ in range(6):   # numpy reuses the freed workl block: same address, no registration pending
        a = ab.solve(None, A.n, 3, 12, "LA", tol=1e-8, mxiter=500, resid=r0, registered_op=A)
        b = ab.solve(A, A.n, 3, 12, "LA", tol=1e-8, mxiter=500, resid=r0)
        assert a.nsteps == 0 and b.nsteps == int(b.iparam[8]) > 0
        assert np.abs(a.d - b.d).max() <= RTOL64 * np.abs(b.d).max()
    # an explicit stale registration on an address a non-applicable solve uses is dropped, not kept
    L = ab.lib()
    w = np.zeros(12 * 12 + 8 * 12)
    assert L.ab200_register_csr_op_f64(w.ctypes.data, A.n, A.nnz, A.rowptr.data_ptr(), A.col.data_ptr(),
                                       A.val.data_ptr()) == 0
    assert L.ab200_register_csr_op_f64(w.ctypes.data, 0, 0, None, None, None) == 0


# --------------------------------------------------------------------------------------------------
# committed golden vectors made by an independent implementation (SciPy's C translation of ARPACK-NG)
# --------------------------------------------------------------------------------------------------
import golden_cases  # noqa: E402


@pytest.mark.parametrize("registered", [False, True])
@pytest.mark.parametrize("c", golden_cases.load(), ids=golden_cases.case_id)
def test_cuda_path_reproduces_committed_scipy_arpack_vectors(ab, c, registered):
    """tests/golden/scipy_arpack_cases.json: same start vector -> same nconv, restart count, OP*x count and
    eigenvalues (1e-10) as SciPy's _arpacklib, through the C-ABI on the device (hand-off loop and registered operator)."""
    S = golden_cases.PROBLEMS[c["problem"]]()
    A = ab.CsrOperator.from_scipy(S)
    r = ab.solve(None if registered else A, A.n, c["nev"], c["ncv"], c["which"], sym=c["sym"], tol=c["tol"],
                 mxiter=3000, resid=golden_cases.start_vector(c, A.n), registered_op=A if registered else None)
    golden_cases.check_against_golden(c, r, c["nev"])


# --------------------------------------------------------------------------------------------------
# round 2: device-resident sweeps, compatibility switch, start-vector rescue
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sym", [True, False])
def test_device_resident_sweeps_take_the_oracles_path_with_one_round_trip_per_sweep(ab, sym):
    """IrlBase::extend, deferred mode on the GPU: same counts and eigenvalues as the oracle, and far fewer blocking
    mailbox reads than Lanczos steps (ab200_host_round_trips)."""
    if sym:
        A, S, nev, ncv, which = ab.CsrOperator.laplace2d(61, 47), laplace2d(61, 47), 5, 24, "LA"
    else:
        A, S, nev, ncv, which = ab.CsrOperator.convdiff2d(40, rho=10.0), convdiff2d(40, rho=10.0), 4, 20, "LM"
    n = A.n
    r0 = np.random.default_rng(31).uniform(-1, 1, n)
    rt0 = ab.host_round_trips()
    g = ab.solve(A, n, nev, ncv, which, sym=sym, tol=1e-10, mxiter=3000, resid=r0, eupd=False)
    trips = ab.host_round_trips() - rt0
    o = Oracle().solve(lambda x: S @ x, n, nev, ncv, which, sym=sym, tol=1e-10, mxiter=3000, resid=r0, c_abi_tol=True,
                       eupd=False)
    assert g.info == o.info == 0 and _counts(g) == _counts(o)
    nopx, sweeps = int(g.iparam[8]), int(g.iparam[2]) + 1
    assert trips <= 3 * sweeps + 6 and trips < nopx // 3


def test_compat_switch_maintains_the_bx_slot_in_mode_1(ab):
    """ab200_set_compat(1): workd(ipntr(3)) = B*x = x at every ido = 1 hand-off, as dsaitr.f:517 leaves it; path and
    results are unchanged."""
    torch = _torch()
    A = ab.CsrOperator.laplace2d(33, 29)
    n = A.n
    r0 = np.random.default_rng(8).uniform(-1, 1, n)
    seen = []

    def op(x, y, *_):
        A(x, y)
        seen.append((x.clone(), None))
    base = ab.solve(A, n, 4, 16, "LA", tol=1e-10, mxiter=500, resid=r0)
    L = ab.lib()
    L.ab200_set_compat(1)
    try:
        checks = []
        v, workd, res = ab.alloc_device_buffers(n, 16)

        def op2(x, y, *_):
            A(x, y)
            checks.append(bool(torch.equal(workd[:n], x)))   # ipntr(3) = 1 in mode 1: the first slot of workd
        r = ab.solve(op2, n, 4, 16, "LA", tol=1e-10, mxiter=500, resid=r0, buffers=(v, workd, res))
    finally:
        L.ab200_set_compat(0)
    assert r.info == 0 and _counts(r) == _counts(base)
    assert np.array_equal(r.d, base.d)
    assert len(checks) > 10 and all(checks[1:])               # (the first product is dgetv0's, ido = -1)


@pytest.mark.parametrize("scale", [1e200, 1e-200])
def test_start_vector_whose_squares_overflow_or_underflow_gpu(ab, scale):
    """IrlBase::rescale_start_vector through the CUDA kernels (k_absmax + exact power-of-two scaling): the oracle's
    path (pdnorm2.f:72-80 / dnrm2 scale by the largest entry)."""
    A, S = ab.CsrOperator.laplace2d(31, 23), laplace2d(31, 23)
    n = A.n
    r0 = np.random.default_rng(77).uniform(-1, 1, n) * scale
    g = ab.solve(A, n, 4, 14, "LA", tol=1e-10, mxiter=500, resid=r0)
    o = Oracle().solve(lambda x: S @ x, n, 4, 14, "LA", tol=1e-10, mxiter=500, resid=r0, c_abi_tol=True)
    assert g.info == o.info == 0 and _counts(g) == _counts(o)
    assert np.abs(g.d - o.d).max() <= RTOL64 * np.abs(o.d).max()
