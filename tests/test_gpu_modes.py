"""GPU parity for the hand-off protocols beyond mode 1: bmat='G' (ido=2), shift-invert,
buckling and Cayley (modes 2-5) for dsaupd/dseupd, real shift-invert for dnaupd/dneupd, 'BE', rvec=0 and the error
exits -- through the C-ABI on the device, against the CPU oracle driven with the same operators and start vector.
The inner solves are the caller's business: they run on the host here (sparse LU), the vectors cross over PCIe."""
import numpy as np
import pytest
import scipy.linalg as sl
import scipy.sparse as sp
import scipy.sparse.linalg as sla

from backends import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ab():
    import arpack_ng_b200 as m
    m.lib()
    return m


def _counts(r):
    return int(r.nconv), int(r.iparam[2]), int(r.iparam[8]), int(r.iparam[9]), int(r.iparam[10])


def _dev(f):
    """Wrap a host function x -> y as a device hand-off: the operand is a CUDA tensor view of workd."""
    import torch

    def g(x, y, *extra):
        xs = x.cpu().numpy() if hasattr(x, "cpu") else x
        out = f(xs, *[e.cpu().numpy() if hasattr(e, "cpu") else e for e in extra])
        if isinstance(y, np.ndarray):
            y[:] = out
        else:
            y.copy_(torch.from_numpy(np.ascontiguousarray(out)))
    return g


def _problem(n=90):
    A = sp.diags([-np.ones(n - 1), 2.2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]).tocsc()
    KG = sp.diags([0.3 * np.ones(n - 1), np.linspace(-1.0, 2.0, n), 0.3 * np.ones(n - 1)], [-1, 0, 1]).tocsc()
    M = sp.diags([np.ones(n - 1) / 6, 4 * np.ones(n) / 6, np.ones(n - 1) / 6], [-1, 0, 1]).tocsc()
    return A, KG, M


@pytest.mark.parametrize("host_buffers", [False, True])
@pytest.mark.parametrize("case", ["mode3_I", "mode3_G", "mode2_G", "mode4_buckling", "mode5_cayley"])
def test_symmetric_modes_vs_oracle(ab, case, host_buffers):
    import torch
    n = 90
    A, KG, M = _problem(n)
    r0 = np.random.default_rng(8).uniform(-1, 1, n)
    kw = dict(tol=1e-12, mxiter=500, resid=r0)
    if case == "mode3_I":
        sigma = 0.1
        lu = sla.splu((A - sigma * sp.identity(n)).tocsc())
        o_op, bop, o_kw = (lambda x: lu.solve(x)), None, dict(mode=3, bmat="I", sigma=sigma)
        d_op = _dev(lambda x: lu.solve(x))
        truth = np.sort(np.linalg.eigvalsh(A.toarray()))
        want = truth[np.argsort(np.abs(truth - sigma))[:4]]
    elif case == "mode3_G":
        sigma = 0.5
        lu = sla.splu((A - sigma * M).tocsc())
        o_op = lambda x, is_bx=False: lu.solve(x if is_bx else M @ x)
        d_op = _dev(lambda x, is_bx=False: lu.solve(x if is_bx else M @ x))
        bop, o_kw = (lambda x: M @ x), dict(mode=3, bmat="G", sigma=sigma)
        truth = np.sort(sl.eigh(A.toarray(), M.toarray(), eigvals_only=True))
        want = truth[np.argsort(np.abs(truth - sigma))[:4]]
    elif case == "mode2_G":
        luM = sla.splu(M)
        o_op = lambda x: (luM.solve(A @ x), A @ x)

        def d_op(x, y, *_):
            xs = x.cpu().numpy() if hasattr(x, "cpu") else x
            ax = A @ xs
            out = luM.solve(ax)
            if isinstance(y, np.ndarray):
                y[:] = out
                x[:] = ax                       # mode 2: x is overwritten with A x (dsaupd.f:309-313)
            else:
                y.copy_(torch.from_numpy(out))
                x.copy_(torch.from_numpy(ax))
        bop, o_kw = (lambda x: M @ x), dict(mode=2, bmat="G", sigma=0.0)
        truth = np.sort(sl.eigh(A.toarray(), M.toarray(), eigvals_only=True))
        want = truth[-4:]
    elif case == "mode4_buckling":
        sigma = 0.6
        lu = sla.splu((A - sigma * KG).tocsc())
        o_op = lambda x, is_bx=False: lu.solve(x if is_bx else A @ x)
        d_op = _dev(lambda x, is_bx=False: lu.solve(x if is_bx else A @ x))
        bop, o_kw = (lambda x: A @ x), dict(mode=4, bmat="G", sigma=sigma)
        lam = sl.eig(A.toarray(), KG.toarray(), right=False)
        truth = np.sort(lam[np.isfinite(lam)].real)
        want = truth[np.argsort(-np.abs(truth / (truth - sigma)))[:4]]   # 'LM' in theta = lambda/(lambda-sigma)
    else:
        sigma = 1.5
        lu = sla.splu((A - sigma * M).tocsc())
        o_op = lambda x, bx: lu.solve(A @ x + sigma * (bx if bx is not None else M @ x))
        d_op = _dev(lambda x, bx=None: lu.solve(A @ x + sigma * (bx if bx is not None else M @ x)))
        bop, o_kw = (lambda x: M @ x), dict(mode=5, bmat="G", sigma=sigma)
        truth = np.sort(sl.eigh(A.toarray(), M.toarray(), eigvals_only=True))
        want = truth[np.argsort(-np.abs((truth + sigma) / (truth - sigma)))[:4]]   # 'LM' in the Cayley theta
    ref = Oracle().solve(o_op, n, 4, 16, "LM", bop=bop, c_abi_tol=True, **o_kw, **kw)
    got = ab.solve(d_op, n, 4, 16, "LM", bop=_dev(bop) if bop else None, host_buffers=host_buffers, **o_kw, **kw)
    assert ref.info == 0 and got.info == 0 and got.ierr == 0
    assert _counts(got) == _counts(ref)
    assert np.abs(np.sort(got.d) - np.sort(ref.d)).max() <= 1e-10 * np.abs(ref.d).max()
    assert np.abs(np.sort(got.d) - np.sort(want)).max() < 1e-7


def test_nonsymmetric_shift_invert_vs_oracle(ab):
    """dnaupd mode 3 with a real shift (dnaupd.f:119-131; TESTS/bug_1323.f is the symmetric twin): OP = inv(A - sigma I)."""
    n = 100
    A = sp.diags([-1.4 * np.ones(n - 1), 2.0 + 0.02 * np.arange(n), -0.5 * np.ones(n - 1)], [-1, 0, 1]).tocsc()
    sigma = 1.0
    lu = sla.splu((A - sigma * sp.identity(n)).tocsc())
    r0 = np.random.default_rng(3).uniform(-1, 1, n)
    kw = dict(sym=False, tol=1e-12, mxiter=1000, mode=3, bmat="I", sigma=sigma, resid=r0)
    ref = Oracle().solve(lambda x: lu.solve(x), n, 4, 20, "LM", c_abi_tol=True, **kw)
    got = ab.solve(_dev(lambda x: lu.solve(x)), n, 4, 20, "LM", **kw)
    assert ref.info == 0 and got.info == 0 and got.ierr == 0
    assert _counts(got) == _counts(ref)
    ev = np.linalg.eigvals(A.toarray())
    lam = got.dr[:got.nconv] + 1j * got.di[:got.nconv]
    for v in lam:
        assert np.abs(ev - v).min() < 1e-6   # the dense eigenvalues of this non-normal band are good to ~1e-8
    assert np.abs(np.sort_complex(lam) - np.sort_complex(ref.dr[:ref.nconv] + 1j * ref.di[:ref.nconv])).max() < 1e-9


def test_both_ends_and_no_vectors(ab):
    """'BE' (dsaup2.f:536-595 exit ordering) and rvec = 0 (values only, dseupd.f:630-648)."""
    A = ab.CsrOperator.laplace2d(31, 24)
    S = A.to_scipy()
    r0 = np.random.default_rng(1).uniform(-1, 1, A.n)
    ref = Oracle().solve(lambda x: S @ x, A.n, 6, 20, "BE", tol=1e-10, mxiter=2000, resid=r0, c_abi_tol=True)
    got = ab.solve(A, A.n, 6, 20, "BE", tol=1e-10, mxiter=2000, resid=r0)
    assert got.info == ref.info == 0 and _counts(got) == _counts(ref)
    assert np.abs(got.d - ref.d).max() <= 1e-10 * np.abs(ref.d).max()
    ev = np.linalg.eigvalsh(S.toarray())
    assert np.abs(np.sort(got.d) - np.concatenate([ev[:3], ev[-3:]])).max() < 1e-8
    nov = ab.solve(A, A.n, 6, 20, "BE", tol=1e-10, mxiter=2000, resid=r0, rvec=False)
    assert nov.info == 0 and nov.ierr == 0
    assert np.abs(np.sort(nov.d) - np.sort(got.d)).max() <= 1e-12 * np.abs(got.d).max()


def test_error_exits_on_device(ab):
    """info = -9 for a zero start vector (dsaup2.f:338-346 via dgetv0), 1 when the budget runs out, 3 never with exact
    shifts; argument errors as in dsaupd.f:493-527 -- all through the device path, none aborts."""
    A = ab.CsrOperator.laplace2d(20, 20)
    z = ab.solve(A, A.n, 3, 12, "LA", tol=1e-10, mxiter=100, resid=np.zeros(A.n), eupd=False)
    assert z.info == -9
    b = ab.solve(A, A.n, 3, 8, "SA", tol=1e-14, mxiter=1, resid=np.ones(A.n), eupd=False)
    S = A.to_scipy()
    ob = Oracle().solve(lambda x: S @ x, A.n, 3, 8, "SA", tol=1e-14, mxiter=1, resid=np.ones(A.n), eupd=False,
                        c_abi_tol=True)
    assert b.info == ob.info == 1 and _counts(b) == _counts(ob)
    for kw, code in ((dict(nev=0), -2), (dict(ncv=3), -3), (dict(which="XY"), -5), (dict(bmat="Q"), -6),
                     (dict(mode=7), -10), (dict(mode=1, bmat="G"), -11)):
        args = dict(nev=3, ncv=12, which="LA", bmat="I", mode=1)
        args.update(kw)
        r = ab.solve(A, A.n, args["nev"], args["ncv"], args["which"], tol=1e-10, mxiter=10, bmat=args["bmat"],
                     mode=args["mode"], bop=lambda x, y: y.copy_(x), eupd=False)
        assert r.info == code, (kw, r.info)
