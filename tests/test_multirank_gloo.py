"""World-size-2 tests of the N > 1 host path on CPU (gloo): the product's host control code in PARPACK mode with its
all-reduces carried by torch.distributed, against the known answer of PARPACK/TESTS/MPI/icb_parpack_c.c and against the
oracle's PARPACK mode; plus the partition helper bench.py shards with."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _rank_main(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist
    from backends import HostDouble, Oracle
    import arpack_ng_b200 as ab
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def allreduce(arr, op):
        t = torch.from_numpy(np.ascontiguousarray(arr))
        dist.all_reduce(t, op=(dist.ReduceOp.SUM, dist.ReduceOp.MAX, dist.ReduceOp.MIN)[op])
        return t.numpy()
    N = 1000
    first, cnt = ab.slab_partition(N, world, rank)
    diag = np.arange(first + 1, first + cnt + 1, dtype=float)
    out = {}
    for name, cls in (("product_host_logic", HostDouble), ("oracle", Oracle)):
        r = cls(rank=rank, nranks=world, allreduce=allreduce).solve(lambda x: diag * x, cnt, 9, 19, "LM", tol=1e-6,
                                                                     mxiter=10000, c_abi_tol=True)
        out[name] = (r.info, r.ierr, r.d.copy(), [int(v) for v in r.iparam], float(np.sum(r.z[:9] ** 2)))
    # the complex twins pznaupd/pzneupd (PARPACK/TESTS/MPI/icb_parpack_c.c:104-190): diag((i+1)(1+i)), rvec = 0
    zdiag = np.arange(first + 1, first + cnt + 1) * (1 + 1j)
    for name, cls in (("product_host_logic_z", HostDouble), ("oracle_z", Oracle)):
        r = cls(rank=rank, nranks=world, allreduce=allreduce).solve_complex(lambda x: zdiag * x, cnt, 9, 19, "LM", tol=1e-6,
                                                                            mxiter=10000, rvec=False, c_abi_tol=True)
        out[name] = (r.info, r.ierr, r.d.copy(), [int(v) for v in r.iparam])
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, out))


def test_pdsaupd_semantics_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=400) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    for rank in (0, 1):
        for name in ("product_host_logic", "oracle"):
            info, ierr, d, iparam, z2 = res[rank][name]
            assert info == 0 and ierr == 0
            assert np.abs(d - np.arange(992, 1001)).max() < 1e-5  # icb_parpack_c.c:93-100
        # host logic == oracle on every rank: counts identical, eigenvalues to rounding
        assert res[rank]["product_host_logic"][3] == res[rank]["oracle"][3]
        assert np.abs(res[rank]["product_host_logic"][2] - res[rank]["oracle"][2]).max() < 1e-9
        # complex twins: known answer (992+i)(1+i), host logic == oracle
        want = np.arange(992, 1001) * (1 + 1j)
        for name in ("product_host_logic_z", "oracle_z"):
            info, ierr, d, iparam = res[rank][name]
            assert info == 0 and ierr == 0
            assert np.abs(d.real - want.real).max() < 1e-5 and np.abs(d.imag - want.imag).max() < 1e-5
        assert res[rank]["product_host_logic_z"][3] == res[rank]["oracle_z"][3]
    # replicated quantities agree across ranks; local Ritz-vector blocks assemble to 9 unit vectors
    assert res[0]["oracle"][3] == res[1]["oracle"][3]
    assert abs(res[0]["product_host_logic"][4] + res[1]["product_host_logic"][4] - 9.0) < 1e-8


def test_slab_partition_covers_everything():
    sys.path.insert(0, ROOT)
    import arpack_ng_b200 as ab
    for n in (1, 7, 512, 4096, 1000):
        for world in (1, 2, 3, 4, 8):
            parts = [ab.slab_partition(n, world, r) for r in range(world)]
            assert parts[0][0] == 0
            assert sum(c for _, c in parts) == n
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f1 == f0 + c0
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
