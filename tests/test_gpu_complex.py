"""GPU parity tests of the complex Arnoldi path (znaupd_c/zneupd_c, cnaupd_c/cneupd_c): the CUDA
path through the C-ABI against the CPU oracle (oracle/ref_impl_complex.inc) on the same seeded inputs.  Tolerances as for
the real path: eigenvalues 1e-10 relative (complex128) / 1e-4 (complex64), ||A x - lambda x|| small, identical nconv,
restart count and OP*x count."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from backends import Oracle
from problems import complex_convdiff2d, complex_tridiag

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ab():
    import arpack_ng_b200 as m
    m.lib()
    return m


def _torch():
    import torch
    return torch


def _dev_op(A):
    """y = A x for a scipy CSR complex matrix, on the device (torch sparse CSR; the OP belongs to the caller)."""
    torch = _torch()
    At = torch.sparse_csr_tensor(torch.as_tensor(A.indptr.astype(np.int64)), torch.as_tensor(A.indices.astype(np.int64)),
                                 torch.as_tensor(A.data), size=A.shape, dtype=torch.as_tensor(A.data).dtype).cuda()

    def op(x, y, *_):
        y.copy_(torch.mv(At, x))
    return op


def _start(n, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)


def _counts(r):
    return int(r.nconv), int(r.iparam[2]), int(r.iparam[8]), int(r.iparam[10])


@pytest.mark.parametrize("name,which", [("tridiag", "LM"), ("tridiag", "SR"), ("tridiag", "SM"), ("conv2d", "LR"),
                                        ("rot_conv2d", "LI"), ("rot_conv2d", "SI")])
def test_znaupd_device_vs_oracle(ab, name, which):
    # rot_conv2d = i * conv2d: the well-separated real ends of the spectrum become its imaginary ends (LI / SI)
    A = {"tridiag": lambda: complex_tridiag(300), "conv2d": lambda: complex_convdiff2d(24),
         "rot_conv2d": lambda: (1j * complex_convdiff2d(24)).tocsr()}[name]()
    n, nev, ncv = A.shape[0], 4, 20
    r0 = _start(n, 5)
    r = ab.solve_complex(_dev_op(A), n, nev, ncv, which, tol=1e-10, mxiter=3000, resid=r0)
    ref = Oracle().solve_complex(lambda x: A @ x, n, nev, ncv, which, tol=1e-10, mxiter=3000, resid=r0, c_abi_tol=True)
    assert r.info == ref.info == 0 and r.ierr == ref.ierr == 0
    assert _counts(r) == _counts(ref)
    assert np.abs(r.d - ref.d).max() <= 1e-10 * np.abs(ref.d).max()
    Z = r.z.cpu().numpy().reshape(ncv, n)[:nev].T
    res = np.linalg.norm(A @ Z - Z * r.d[None, :], axis=0)
    assert (res <= 1e-8 * np.abs(r.d).max()).all(), res
    assert np.abs(np.linalg.norm(Z, axis=0) - 1).max() < 1e-10
    assert ab.launch_stats()["kernels"] > 0


def test_znaupd_host_buffers_match_device_buffers(ab):
    """An unmodified CPU caller (host resid/v/workd, CPU OP) and the device-pointer caller run the same kernels."""
    A = complex_tridiag(257)
    n, nev, ncv = A.shape[0], 3, 16
    r0 = _start(n, 9)

    def op_host(x, y, *_):
        y[:] = A @ x
    rd = ab.solve_complex(_dev_op(A), n, nev, ncv, "LM", tol=1e-10, mxiter=3000, resid=r0)
    rh = ab.solve_complex(op_host, n, nev, ncv, "LM", tol=1e-10, mxiter=3000, resid=r0, host_buffers=True)
    assert rd.info == rh.info == 0 and rd.ierr == rh.ierr == 0
    assert _counts(rd)[:3] == _counts(rh)[:3]
    assert np.abs(rd.d - rh.d).max() <= 1e-10 * np.abs(rd.d).max()
    Z = rh.z.reshape(ncv, n)[:nev].T
    assert (np.linalg.norm(A @ Z - Z * rh.d[None, :], axis=0) <= 1e-8 * np.abs(rh.d).max()).all()


def test_cnaupd_single_precision(ab):
    A = complex_tridiag(200, imag_span=20.0).astype(np.complex64)
    n, nev, ncv = A.shape[0], 3, 16
    r0 = _start(n, 2).astype(np.complex64)
    r = ab.solve_complex(_dev_op(A), n, nev, ncv, "LM", tol=1e-5, mxiter=3000, resid=r0, dtype=np.complex64)
    ref = Oracle().solve_complex(lambda x: A @ x, n, nev, ncv, "LM", tol=1e-5, mxiter=3000, resid=r0,
                                 dtype=np.complex64, c_abi_tol=True)
    assert r.info == ref.info == 0 and r.ierr == 0
    assert r.nconv == ref.nconv == nev
    assert np.abs(np.sort_complex(r.d) - np.sort_complex(ref.d)).max() <= 1e-4 * np.abs(ref.d).max()


def test_znaupd_shift_invert_mode3(ab):
    """Mode 3 (zndrv2-style): OP = inv(A - sigma I), eigenvalues of A nearest sigma; exercises the back-transform and
    the zgeru purification of zneupd (zneupd.f:820-868).  The caller's OP is a CPU sparse LU, so host buffers."""
    A = complex_tridiag(220)
    n, nev, ncv = A.shape[0], 4, 20
    sigma = 50000.0 + 20.0j
    lu = spla.splu((A - sigma * __import__("scipy.sparse").sparse.eye(n)).tocsc())
    r0 = _start(n, 4)

    def op_host(x, y, *_):
        y[:] = lu.solve(np.ascontiguousarray(x))
    r = ab.solve_complex(op_host, n, nev, ncv, "LM", tol=1e-10, mxiter=3000, resid=r0, mode=3, sigma=sigma,
                         host_buffers=True)
    ref = Oracle().solve_complex(lambda x: lu.solve(np.ascontiguousarray(x)), n, nev, ncv, "LM", tol=1e-10, mxiter=3000,
                                 resid=r0, mode=3, sigma=sigma, c_abi_tol=True)
    assert r.info == ref.info == 0 and r.ierr == ref.ierr == 0
    assert _counts(r)[:3] == _counts(ref)[:3]
    assert np.abs(r.d - ref.d).max() <= 1e-10 * np.abs(ref.d).max()
    Z = r.z.reshape(ncv, n)[:nev].T
    assert (np.linalg.norm(A @ Z - Z * r.d[None, :], axis=0) <= 1e-7 * np.abs(r.d).max()).all()
    dense = np.linalg.eigvals(A.toarray())
    want = dense[np.argsort(np.abs(dense - sigma))[:nev]]
    assert np.abs(np.sort_complex(r.d) - np.sort_complex(want)).max() <= 1e-8 * np.abs(want).max()


def test_znaupd_argument_errors_and_no_vectors(ab):
    A = complex_tridiag(64)
    n = A.shape[0]
    r = ab.solve_complex(_dev_op(A), n, 3, 3, "LM", eupd=False)          # ncv <= nev
    assert r.info == -3
    r = ab.solve_complex(_dev_op(A), n, 3, 12, "LA", eupd=False)         # which of the symmetric family
    assert r.info == -5
    r0 = _start(n, 1)
    a = ab.solve_complex(_dev_op(A), n, 3, 12, "LM", tol=1e-10, mxiter=3000, resid=r0)
    b = ab.solve_complex(_dev_op(A), n, 3, 12, "LM", tol=1e-10, mxiter=3000, resid=r0, rvec=False)
    assert a.info == b.info == 0 and b.ierr == 0
    assert np.abs(np.sort_complex(a.d) - np.sort_complex(b.d)).max() <= 1e-10 * np.abs(a.d).max()


def test_icb_arpack_c_zn(ab):
    """TESTS/icb_arpack_c.c:98-165 through the C-ABI on the device: A = diag((i+1)(1+i)), nev=9, ncv=19, 'LM',
    tol=1e-6, rvec=0, random start (zlarnv stream) -> d[i] = (992+i)(1+i), 1e-5 per component."""
    torch = _torch()
    n, nev, ncv = 1000, 9, 19
    diag = torch.arange(1, n + 1, dtype=torch.float64, device="cuda") * (1 + 1j)
    ab.lib().ab200_reset_seed()
    r = ab.solve_complex(lambda x, y, *_: torch.mul(diag, x, out=y), n, nev, ncv, "LM", tol=1e-6, mxiter=10 * n,
                         rvec=False)
    assert r.info == 0 and r.ierr == 0 and r.nconv >= nev
    want = (n - (nev - 1) + np.arange(nev)) * (1 + 1j)
    assert np.abs(r.d.real - want.real).max() <= 1e-5 and np.abs(r.d.imag - want.imag).max() <= 1e-5
    ref = Oracle().solve_complex(lambda x: np.arange(1, n + 1) * (1 + 1j) * x, n, nev, ncv, "LM", tol=1e-6,
                                 mxiter=10 * n, rvec=False, c_abi_tol=True)
    assert _counts(r)[:3] == _counts(ref)[:3]     # same zlarnv start vector, same path


import golden_cases  # noqa: E402


@pytest.mark.parametrize("c", golden_cases.load_complex(), ids=golden_cases.case_id)
def test_cuda_path_reproduces_committed_scipy_znaupd_vectors(ab, c):
    """Golden vectors made by an implementation that is not ours (SciPy's C translation of znaupd/zneupd): identical
    nconv, restart and OP*x counts, eigenvalues to 1e-10, through the C-ABI on the device."""
    A = golden_cases.ZPROBLEMS[c["problem"]]()
    n = A.shape[0]
    r = ab.solve_complex(_dev_op(A), n, c["nev"], c["ncv"], c["which"], tol=c["tol"], mxiter=3000,
                         resid=golden_cases.start_vector_complex(c, n))
    golden_cases.check_against_golden_complex(c, r)


# --------------------------------------------------------------------------------------------------
# per-kernel tests: each complex kernel against a plain numpy statement of the same op
# --------------------------------------------------------------------------------------------------
def _crand(rng, *shape):
    return rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)


@pytest.mark.parametrize("n,j,pad", [(1, 1, 0), (37, 5, 0), (1000, 8, 3), (1000, 9, 0), (100003, 31, 1),
                                     (65536, 64, 0), (4099, 70, 5)])
def test_kernel_zdots_zupdate_vs_numpy(ab, n, j, pad):
    """k_zdots (h = V_j^H w, sum conj(w) w) and k_zupdate (r = w - V_j h, sum |r|^2) on odd sizes, column counts around
    the 8-column register chunk, padded leading dimensions."""
    torch = _torch()
    rng = np.random.default_rng(n + j)
    ldv = n + pad
    V = np.zeros((j, ldv), dtype=complex)       # column-major (ldv, j)
    V[:, :n] = _crand(rng, j, n) / np.sqrt(n)
    w = _crand(rng, n)
    Vd, wd = torch.as_tensor(V.ravel()).cuda(), torch.as_tensor(w).cuda()
    rd = torch.zeros(n, dtype=torch.complex128, device="cuda")
    out = np.zeros(j + 2, dtype=complex)
    assert ab.lib().ab200_debug_zorth_f64(n, j, Vd.data_ptr(), ldv, wd.data_ptr(), rd.data_ptr(), out.ctypes.data) == 0
    Vn = V[:, :n].T                              # n x j
    h = Vn.conj().T @ w
    r = w - Vn @ h
    scale = np.abs(h).max() + 1e-300
    assert np.abs(out[:j] - h).max() <= 1e-12 * max(scale, 1.0)
    assert abs(out[j] - np.vdot(w, w)) <= 1e-12 * abs(np.vdot(w, w)) and abs(out[j].imag) <= 1e-12 * abs(out[j].real)
    # the device used ITS h for the update: compare against numpy's r up to that difference
    assert np.abs(rd.cpu().numpy() - r).max() <= 1e-11 * max(1.0, np.abs(w).max())
    assert abs(out[j + 1].real - np.vdot(r, r).real) <= 1e-10 * np.vdot(r, r).real and out[j + 1].imag == 0.0


@pytest.mark.parametrize("n,kin,kout,beta_col,inplace", [(1, 3, 2, 1, True), (1000, 20, 7, 6, True), (1001, 20, 8, -1, True),
                                                         (70000, 30, 30, -1, True), (4097, 64, 21, 20, True),
                                                         (513, 12, 5, -1, False), (3000, 210, 9, 8, True)])
def test_kernel_zvq_vs_numpy(ab, n, kin, kout, beta_col, inplace):
    """k_zvq: V(:,0:kout) <- V(:,0:kin) Q in place (or into a separate array), resid <- sigma resid + beta Vnew(:,col)
    and its squared norm; tile sizes 64 and 32 rows (kin > 200), column counts that are not multiples of the groups."""
    torch = _torch()
    rng = np.random.default_rng(n + kin + kout)
    ldv = n + 2
    V = np.zeros((kin, ldv), dtype=complex)
    V[:, :n] = _crand(rng, kin, n)
    Q = _crand(rng, kout, kin)                   # column-major kin x kout
    resid = _crand(rng, n)
    sigma, beta = complex(0.3, -0.7), complex(-1.1, 0.4)
    Vd = torch.as_tensor(V.ravel()).cuda()
    rd = torch.as_tensor(resid).cuda()
    nrm = np.zeros(1)
    want = V[:, :n].T @ Q.T                      # n x kout
    L = ab.lib()
    if inplace:
        assert L.ab200_debug_zvq_f64(n, kin, kout, Vd.data_ptr(), ldv, Q.ctypes.data, Vd.data_ptr(), ldv, sigma.real,
                                     sigma.imag, beta.real, beta.imag, beta_col, rd.data_ptr(),
                                     nrm.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_double))) == 0
        got = Vd.cpu().numpy().reshape(kin, ldv)[:kout, :n].T
        rwant = sigma * resid + (beta * want[:, beta_col] if beta_col >= 0 else 0.0)
        assert np.abs(rd.cpu().numpy() - rwant).max() <= 1e-11 * max(1.0, np.abs(rwant).max())
        assert abs(nrm[0] - np.vdot(rwant, rwant).real) <= 1e-10 * np.vdot(rwant, rwant).real
        # columns kout..kin of V are untouched
        assert np.array_equal(Vd.cpu().numpy().reshape(kin, ldv)[kout:, :n], V[kout:, :n])
    else:
        Od = torch.zeros(kout * n, dtype=torch.complex128, device="cuda")
        assert L.ab200_debug_zvq_f64(n, kin, kout, Vd.data_ptr(), ldv, Q.ctypes.data, Od.data_ptr(), n, 0.0, 0.0, 0.0, 0.0,
                                     -1, None, nrm.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_double))) == 0
        got = Od.cpu().numpy().reshape(kout, n).T
        assert np.array_equal(Vd.cpu().numpy(), V.ravel())
    assert np.abs(got - want).max() <= 1e-11 * max(1.0, np.abs(want).max())
