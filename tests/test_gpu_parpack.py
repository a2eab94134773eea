"""PARPACK entry points on the GPU (pdsaupd_c/pdseupd_c/pdnaupd_c/pdneupd_c with an NCCL communicator handle).
World size 1 runs in the ordinary single-GPU suite (a 1-rank NCCL communicator: the PARPACK semantics -- per-rank
seed, no initial OP*x, eps23 exponent -- and the all-reduce plumbing are exercised against the oracle's PARPACK mode);
with >= 2 visible GPUs tools/multigpu_check.py is run under torchrun as well."""
import os
import subprocess
import sys

import numpy as np
import pytest

from backends import Oracle
from problems import convdiff2d, laplace2d

pytestmark = pytest.mark.gpu

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _self_allreduce(arr, op):
    return arr


def _counts(r):
    return int(r.nconv), int(r.iparam[2]), int(r.iparam[8]), int(r.iparam[10])


def test_icb_parpack_c_single_rank(ab_comm):
    """PARPACK/TESTS/MPI/icb_parpack_c.c:30-102 with one rank: diag(1..1000) -> 992..1000 (1e-5), random start from
    pdgetv0's per-rank seed; counts identical to the oracle's PARPACK mode."""
    import torch
    ab, comm = ab_comm
    n = 1000
    diag = torch.arange(1, n + 1, dtype=torch.float64, device="cuda")
    ab.lib().ab200_reset_seed()
    r = ab.solve(lambda x, y, *_: torch.mul(diag, x, out=y), n, 9, 19, "LM", tol=1e-6, mxiter=10000, comm=comm)
    assert r.info == 0 and r.ierr == 0
    assert np.abs(r.d - np.arange(992, 1001)).max() < 1e-5
    ref = Oracle(rank=0, nranks=1, allreduce=_self_allreduce).solve(lambda x: np.arange(1, n + 1) * x, n, 9, 19, "LM",
                                                                    tol=1e-6, mxiter=10000, c_abi_tol=True)
    assert _counts(r) == _counts(ref)
    assert ab.lib().ab200_comm_rank(comm) == 0 and ab.lib().ab200_comm_size(comm) == 1


@pytest.mark.parametrize("sym", [True, False])
def test_parpack_semantics_differ_from_serial_as_in_the_reference(ab_comm, sym):
    """Appendix B.11: pdgetv0 spends no OP*x on the start vector for bmat='I' (pdgetv0.f:285), dgetv0 always does
    (dgetv0.f:245-251) -- so for the same resid the two entry points take different paths.  Each must match its own
    oracle mode exactly."""
    ab, comm = ab_comm
    if sym:
        A, nev, ncv, which = ab.CsrOperator.laplace2d(33, 27), 4, 16, "LA"
        S = laplace2d(33, 27)
    else:
        A, nev, ncv, which = ab.CsrOperator.convdiff2d(30, rho=10.0), 4, 20, "LM"
        S = convdiff2d(30, rho=10.0)
    n = A.n
    r0 = np.random.default_rng(9).uniform(-1, 1, n)
    par = ab.solve(A, n, nev, ncv, which, sym=sym, tol=1e-10, mxiter=3000, resid=r0, comm=comm)
    ser = ab.solve(A, n, nev, ncv, which, sym=sym, tol=1e-10, mxiter=3000, resid=r0)
    opar = Oracle(rank=0, nranks=1, allreduce=_self_allreduce).solve(lambda x: S @ x, n, nev, ncv, which, sym=sym,
                                                                     tol=1e-10, mxiter=3000, resid=r0, c_abi_tol=True)
    oser = Oracle().solve(lambda x: S @ x, n, nev, ncv, which, sym=sym, tol=1e-10, mxiter=3000, resid=r0, c_abi_tol=True)
    assert par.info == ser.info == opar.info == oser.info == 0
    assert _counts(par) == _counts(opar)
    assert _counts(ser) == _counts(oser)
    assert int(ser.iparam[8]) != int(par.iparam[8]) or int(ser.iparam[2]) != int(par.iparam[2])
    if sym:
        assert np.abs(par.d - opar.d).max() <= 1e-10 * np.abs(opar.d).max()
    else:
        assert np.abs(np.sort(par.dr[:nev]) - np.sort(opar.dr[:nev])).max() <= 1e-10 * np.abs(opar.dr).max()


def test_registered_operator_under_a_communicator(ab_comm):
    """ab200_register_csr_halo_op_f64: pdsaupd_c applies the (here halo-less, 1 rank) operator itself: one call per
    solve, same counts and eigenvalues as the hand-off loop."""
    ab, comm = ab_comm
    A = ab.CsrOperator.laplace2d(41, 37)
    r0 = np.random.default_rng(4).uniform(-1, 1, A.n)
    a = ab.solve(A, A.n, 5, 20, "LA", tol=1e-10, mxiter=2000, resid=r0, comm=comm)
    b = ab.solve(None, A.n, 5, 20, "LA", tol=1e-10, mxiter=2000, resid=r0, comm=comm, registered_op=A)
    assert a.info == b.info == 0 and a.ierr == b.ierr == 0
    assert b.nsteps == 0 and a.nsteps == int(a.iparam[8])
    assert _counts(a) == _counts(b)
    assert np.abs(a.d - b.d).max() <= 1e-10 * np.abs(a.d).max()
    # a halo registration on the sequential entry point is refused loudly
    with pytest.raises(ab.ArpackB200Error):
        L = ab.lib()
        w = None
        import torch
        halo = torch.zeros(1, dtype=torch.float64, device="cuda")

        def hijack():
            res = ab.alloc_device_buffers(A.n, 20)
            workl = np.zeros(20 * 20 + 8 * 20)
            assert L.ab200_register_csr_halo_op_f64(workl.ctypes.data, comm, A.n, A.nnz, A.rowptr.data_ptr(),
                                                    A.col.data_ptr(), A.val.data_ptr(), 0, 0, halo.data_ptr()) == 0
            ido, info = np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.int32)
            iparam, ipntr = np.zeros(11, dtype=np.int32), np.zeros(14, dtype=np.int32)
            iparam[[0, 2, 3, 6]] = [1, 10, 1, 1]
            ab.dsaupd_c(ido, "I", A.n, "LA", 5, 1e-10, res[2], 20, res[0], A.n, iparam, ipntr, res[1], workl, info)
        hijack()


def test_pdseupd_rvec0_shift_invert_does_not_touch_z(ab_comm):
    """pdseupd_c with rvec = 0 in mode 3: eigenvalues only.  The reference's pdseupd.f:858 runs its rank-1
    purification update even then; here z is not mapped without rvec, so the update must be skipped (it used to
    write through a null device pointer and poison the CUDA context).  Values equal the rvec = 1 run's."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    ab, comm = ab_comm
    S = laplace2d(19, 17).tocsc()
    n, sigma = S.shape[0], 0.9
    lu = spla.splu((S - sigma * sp.identity(n)).tocsc())
    r0 = np.random.default_rng(12).uniform(-1, 1, n)

    def op(x, y, *_):
        y[:] = lu.solve(np.ascontiguousarray(x))
    kw = dict(tol=1e-10, mxiter=2000, mode=3, sigma=sigma, resid=r0, comm=comm, host_buffers=True)
    a = ab.solve(op, n, 4, 16, "LM", rvec=True, **kw)
    b = ab.solve(op, n, 4, 16, "LM", rvec=False, **kw)
    assert a.info == b.info == 0 and a.ierr == b.ierr == 0
    assert np.abs(np.sort(a.d) - np.sort(b.d)).max() <= 1e-12 * np.abs(a.d).max()
    import torch
    torch.cuda.synchronize()     # a faulting k_ger would surface here as a sticky error


def test_multi_gpu_check_under_torchrun_when_available():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 visible GPUs (run by hand: gpurun --gpus 2 -- torchrun ... tools/multigpu_check.py)")
    nproc = 2
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29700 + os.getpid() % 1000),
                        os.path.join(_ROOT, "tools", "multigpu_check.py")], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "PASS" in p.stdout
