"""PARPACK complex twins on the GPU: pznaupd_c/pzneupd_c with a 1-rank NCCL communicator (the PARPACK semantics of
PARPACK/SRC/MPI/pzn*.f -- per-rank zlarnv seed, no initial OP*x for bmat = 'I', REAL eps23 exponent -- and the all-reduce
plumbing of the complex mailbox) against the oracle's PARPACK mode and the known answer of
PARPACK/TESTS/MPI/icb_parpack_c.c:104-190.

The CPU side of the same path (oracle and host logic on 2 and 3 logical ranks) is covered by
tests/test_complex_cpu.py; the kernels are those of the sequential complex path.  Observed passing on a B200 with a
1-rank communicator; a multi-rank GPU run has not been made yet."""
import numpy as np
import pytest

from backends import Oracle

pytestmark = pytest.mark.gpu


def _self_allreduce(arr, op):
    return arr


def test_icb_parpack_c_zn_single_rank(ab_comm):
    import torch
    ab, comm = ab_comm
    n, nev, ncv = 1000, 9, 19
    diag = torch.arange(1, n + 1, dtype=torch.float64, device="cuda") * (1 + 1j)
    ab.lib().ab200_reset_seed()
    r = ab.solve_complex(lambda x, y, *_: torch.mul(diag, x, out=y), n, nev, ncv, "LM", tol=1e-6, mxiter=10 * n,
                         rvec=False, comm=comm)
    assert r.info == 0 and r.ierr == 0 and r.nconv >= nev
    want = (n - (nev - 1) + np.arange(nev)) * (1 + 1j)
    assert np.abs(r.d.real - want.real).max() <= 1e-5 and np.abs(r.d.imag - want.imag).max() <= 1e-5
    ref = Oracle(rank=0, nranks=1, allreduce=_self_allreduce).solve_complex(
        lambda x: np.arange(1, n + 1) * (1 + 1j) * x, n, nev, ncv, "LM", tol=1e-6, mxiter=10 * n, rvec=False,
        c_abi_tol=True)
    assert (r.nconv, int(r.iparam[2]), int(r.iparam[8])) == (ref.nconv, int(ref.iparam[2]), int(ref.iparam[8]))
