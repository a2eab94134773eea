"""Multi-GPU parity against the oracle's logical-rank (PARPACK) mode, run under torchrun (one rank per GPU, NCCL):

every GPU rank r drives the product's p*aupd_c / p*eupd_c on its row block AND, on the host, the oracle restated from
PARPACK/SRC/MPI (Oracle(rank=r, nranks=N, allreduce=gloo)) on the SAME partition with the same operator, start vector
and seeds; the two must take the same path -- identical (nconv, restarts, OP*x, re-orthogonalisations) -- and return
the same eigenvalues to 1e-10 relative (the bar of BASELINE.json's north_star).

  1. pdsaupd_c/pdseupd_c: PARPACK/TESTS/MPI/icb_parpack_c.c:60-102 -- diag(1..1000) split over the ranks, random start
     vector from pdgetv0's per-rank seeds (pdgetv0.f:234-245) -> 992..1000; all-reduce sites pdsaitr.f:604,720
  2. pdsaupd_c: 3-D 7-point Laplacian, z-slab partition with halo exchange (pdsdrv1.f:463-483), nev 6 ncv 24 'LA'
  3. the same solve with the operator registered (ab200_register_csr_halo_op_f64): no hand-off, same path
  4. pdnaupd_c/pdneupd_c: 2-D convection-diffusion (dndrv1.f:453-470, rho = 10), row blocks, nev 4 ncv 20 'LM'
  5. pznaupd_c/pzneupd_c: icb_parpack_c.c:104-190 -- diag((i+1)(1+i)), rvec = 0
  8. pdsaupd_c with a start vector inside an invariant subspace: the device-resident sweep is cut short on every rank
  7. pssaupd_c (FP32): 2-D Laplacian by y-slabs, eigenvalues to 1e-4 (north_star's FP32 bar)
  6. BASELINE config 5 at test size: SVD through pdsaupd_c on A^T A, A 20k x 5k with 16 nnz/row ROW-SHARDED over the
     ranks (all-gather x, local A and A^T products, reduce-scatter; EXAMPLES/SVD/dsvd.f:342-343), hand-off and registered

Run it twice per GPU count: with the peer-memory reductions (default) and with AB200_P2P=0 (ncclAllReduce).
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/multigpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import arpack_ng_b200 as ab  # noqa: E402
from backends import Oracle  # noqa: E402  (the checker; never on the product path)
from problems import convdiff2d, laplace3d  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gloo = dist.new_group(backend="gloo")      # carries the oracle's MPI_ALLREDUCEs between the host processes
comm = ab.nccl_comm_from_torch_distributed()
L = ab.lib()
ok = True
path = "peer-memory" if L.ab200_comm_uses_p2p(comm) else "nccl"


def host_allreduce(arr, op):
    t = torch.from_numpy(np.ascontiguousarray(arr))
    dist.all_reduce(t, op=(dist.ReduceOp.SUM, dist.ReduceOp.MAX, dist.ReduceOp.MIN)[op], group=gloo)
    return t.numpy()


def host_allgather(x, counts):
    """global vector from the ranks' row blocks (host side, for the oracle's OP)."""
    parts = [None] * world
    dist.all_gather_object(parts, np.ascontiguousarray(x), group=gloo)
    return np.concatenate(parts)


def counts_of(r):
    return int(r.nconv), int(r.iparam[2]), int(r.iparam[8]), int(r.iparam[10])


def report(name, good, detail):
    global ok
    ok &= bool(good)
    print(f"[rank {rank}/{world} {path}] {name}: {'ok' if good else 'MISMATCH'} {detail}", flush=True)


def oracle():
    return Oracle(rank=rank, nranks=world, allreduce=host_allreduce)


# ---- 1. icb_parpack_c: pdsaupd_c, random start from the per-rank seeds ----
N = 1000
first, cnt = ab.slab_partition(N, world, rank)
diag_h = np.arange(first + 1, first + cnt + 1, dtype=float)
diag = torch.as_tensor(diag_h, device="cuda")
L.ab200_reset_seed()
g = ab.solve(lambda x, y, *_: torch.mul(diag, x, out=y), cnt, 9, 19, "LM", tol=1e-6, mxiter=10000, comm=comm)
o = oracle().solve(lambda x: diag_h * x, cnt, 9, 19, "LM", tol=1e-6, mxiter=10000, c_abi_tol=True)
err = np.abs(g.d - np.arange(992, 1001)).max()
rel = np.abs(g.d - o.d).max() / np.abs(o.d).max()
report("pdsaupd_c diag(1..1000)", g.info == 0 and g.ierr == 0 and err < 1e-5 and counts_of(g) == counts_of(o) and
       rel <= 1e-10, f"counts gpu={counts_of(g)} oracle={counts_of(o)} |d-(992..1000)|={err:.1e} rel vs oracle={rel:.1e}")

# ---- 2. 3-D Laplacian with halos: pdsaupd_c ----
nx, ny, nz = 24, 20, max(16, 4 * world)
z0, nzloc = ab.slab_partition(nz, world, rank)
A = ab.CsrOperator.laplace3d(nx, ny, nz, z0=z0, nzloc=nzloc)
nloc = A.n
all_counts = [ab.slab_partition(nz, world, r)[1] * nx * ny for r in range(world)]
r0 = ab.hashed_start_vector(nloc, i0=z0 * nx * ny)
r0h = ab.hashed_start_vector_numpy(nloc, i0=z0 * nx * ny)
As = laplace3d(nx, ny, nz).tocsr()
rows = slice(z0 * nx * ny, z0 * nx * ny + nloc)
As_loc = As[rows, :]


def op_host(x):   # the oracle's av: gather the global vector, multiply my rows (what pdsdrv1.f's halo exchange amounts to)
    return As_loc @ host_allgather(x, all_counts)


g = ab.solve(lambda x, y, *_: A.apply_halo(comm, x, y), nloc, 6, 24, "LA", tol=1e-10, mxiter=2000, resid=r0, comm=comm)
o = oracle().solve(op_host, nloc, 6, 24, "LA", tol=1e-10, mxiter=2000, resid=r0h, c_abi_tol=True)
rel = np.abs(g.d - o.d).max() / np.abs(o.d).max()
report(f"pdsaupd_c laplace3d {nx}x{ny}x{nz}", g.info == 0 and g.ierr == 0 and counts_of(g) == counts_of(o) and rel <= 1e-10,
       f"counts gpu={counts_of(g)} oracle={counts_of(o)} rel eig diff={rel:.1e}")
# Ritz vector blocks: global residual || A z - d z || through the halo operator
z = g.z[:nloc * g.nconv].view(g.nconv, nloc)
y = torch.empty(nloc, dtype=torch.float64, device="cuda")
rn = torch.zeros(g.nconv, dtype=torch.float64, device="cuda")
for k in range(g.nconv):
    A.apply_halo(comm, z[k].contiguous(), y)
    rn[k] = torch.sum((y - g.d[k] * z[k]) ** 2)
dist.all_reduce(rn)
rn = torch.sqrt(rn).cpu().numpy()
report("global residuals ||A z - d z||", bool((rn <= 1e-10 * 12.0 * 10).all()), f"max={rn.max():.2e}")

# ---- 3. the same solve with the operator registered (halo exchange + fused SpMV inside pdsaupd_c) ----
reg = ab.solve(None, nloc, 6, 24, "LA", tol=1e-10, mxiter=2000, resid=r0, comm=comm, registered_op=A)
same = (reg.info == 0 and reg.ierr == 0 and reg.nsteps == 0 and counts_of(reg) == counts_of(o) and
        np.abs(reg.d - o.d).max() <= 1e-10 * np.abs(o.d).max())
report("registered halo operator", same, f"counts={counts_of(reg)} hand-offs={reg.nsteps}")

# ---- 4. pdnaupd_c: convection-diffusion, row blocks; the caller's OP gathers x over NCCL ----
m = 31   # 961 rows: the ranks' row counts differ, and some are odd -- those ranks cannot use the TMA-tiled kernels
# (16-byte column stride), so the ranks must AGREE not to fuse their reductions (CudaVecOps::ranks_agree_on_fusing)
S = convdiff2d(m, rho=10.0).tocsr()
nn = m * m
f4, c4 = ab.slab_partition(nn, world, rank)
cnts4 = [ab.slab_partition(nn, world, r)[1] for r in range(world)]
S_loc = S[f4:f4 + c4, :].tocsr()
Sd = ab.CsrOperator.from_scipy(S_loc)
xg = torch.zeros(nn, dtype=torch.float64, device="cuda")
chunks = list(torch.split(xg, cnts4))


def op4(x, yv, *_):
    chunks[rank].copy_(x)
    for r in range(world):          # row blocks may differ in size: one broadcast per owner
        dist.broadcast(chunks[r], src=r)
    Sd(xg, yv)


r4 = np.random.default_rng(11).uniform(-1, 1, nn)[f4:f4 + c4]
g = ab.solve(op4, c4, 4, 20, "LM", sym=False, tol=1e-10, mxiter=3000, resid=r4, comm=comm)
o = oracle().solve(lambda x: S_loc @ host_allgather(x, cnts4), c4, 4, 20, "LM", sym=False, tol=1e-10, mxiter=3000,
                   resid=r4, c_abi_tol=True)
ev_g = np.sort_complex(g.dr[:4] + 1j * g.di[:4])
ev_o = np.sort_complex(o.dr[:4] + 1j * o.di[:4])
rel = np.abs(ev_g - ev_o).max() / np.abs(ev_o).max()
report("pdnaupd_c convdiff2d", g.info == 0 and g.ierr == 0 and counts_of(g) == counts_of(o) and rel <= 1e-10,
       f"counts gpu={counts_of(g)} oracle={counts_of(o)} rel eig diff={rel:.1e}")

# ---- 5. pznaupd_c: icb_parpack_c.c:104-190 ----
zdiag_h = np.arange(first + 1, first + cnt + 1) * (1 + 1j)
zdiag = torch.as_tensor(zdiag_h, device="cuda")
L.ab200_reset_seed()
g = ab.solve_complex(lambda x, y, *_: torch.mul(zdiag, x, out=y), cnt, 9, 19, "LM", tol=1e-6, mxiter=10000, rvec=False,
                     comm=comm)
o = oracle().solve_complex(lambda x: zdiag_h * x, cnt, 9, 19, "LM", tol=1e-6, mxiter=10000, rvec=False, c_abi_tol=True)
want = np.arange(992, 1001) * (1 + 1j)
dz = np.sort_complex(np.asarray(g.d))
err = np.abs(dz - np.sort_complex(want)).max()
rel = np.abs(dz - np.sort_complex(np.asarray(o.d))).max() / np.abs(want).max()
report("pznaupd_c diag((i+1)(1+i))", g.info == 0 and g.ierr == 0 and err < 1e-5 and
       (int(g.nconv), int(g.iparam[2]), int(g.iparam[8])) == (int(o.nconv), int(o.iparam[2]), int(o.iparam[8])) and
       rel <= 1e-10, f"nconv/restarts/nopx gpu={(int(g.nconv), int(g.iparam[2]), int(g.iparam[8]))} "
       f"oracle={(int(o.nconv), int(o.iparam[2]), int(o.iparam[8]))} err={err:.1e} rel vs oracle={rel:.1e}")

# ---- 6. config 5 at test size: row-sharded A^T A under pdsaupd_c ----
M5, K5 = 20000, 5000
G5 = ab.GramOperator.randsparse(M5, K5, 16, comm=comm, shard_rows=1500)
A5 = ab.randsparse_numpy(0, M5, K5, 16)
k_loc = K5 // world
r5 = ab.hashed_start_vector_numpy(k_loc, i0=rank * k_loc)
cnts5 = [k_loc] * world


def op5_host(x):   # the oracle's av + atv: gather x, the full product, keep my slice
    return (A5.T @ (A5 @ host_allgather(x, cnts5)))[rank * k_loc:(rank + 1) * k_loc]


o = oracle().solve(op5_host, k_loc, 16, 48, "LM", tol=1e-10, mxiter=3000, resid=r5, c_abi_tol=True)
for reg in (False, True):
    g = ab.solve(None if reg else G5, k_loc, 16, 48, "LM", tol=1e-10, mxiter=3000, resid=r5, comm=comm,
                 registered_op=G5 if reg else None)
    sg, so = np.sqrt(np.maximum(g.d, 0)), np.sqrt(np.maximum(o.d, 0))
    rel = np.abs(sg - so).max() / so.max()
    report(f"pdsaupd_c on row-sharded A^T A ({'registered' if reg else 'hand-off'})",
           g.info == 0 and g.ierr == 0 and counts_of(g) == counts_of(o) and rel <= 1e-10 and (g.nsteps == 0) == reg,
           f"counts gpu={counts_of(g)} oracle={counts_of(o)} rel sigma diff={rel:.1e} sigma_max={sg.max():.6f}")
G5.close()

# ---- 7. pssaupd_c: the single-precision twin (fused reductions on 4-byte values), 2-D Laplacian by y-slabs ----
nx7, ny7 = 40, 8 * world
y0, nyl = ab.slab_partition(ny7, world, rank)
n7 = nyl * nx7
cnts7 = [ab.slab_partition(ny7, world, r)[1] * nx7 for r in range(world)]
from problems import laplace2d  # noqa: E402
S7 = laplace2d(nx7, ny7).tocsr().astype(np.float32)
S7_loc = S7[y0 * nx7:(y0 + nyl) * nx7, :]
r7 = np.random.default_rng(5).uniform(-1, 1, nx7 * ny7).astype(np.float32)[y0 * nx7:(y0 + nyl) * nx7]
xg7 = torch.zeros(nx7 * ny7, dtype=torch.float32, device="cuda")
ch7 = list(torch.split(xg7, cnts7))
S7d = ab.CsrOperator.from_scipy(S7_loc.astype(np.float64))
S7d.val = S7d.val.to(torch.float32)


def op7(x, yv, *_):
    ch7[rank].copy_(x)
    for r in range(world):
        dist.broadcast(ch7[r], src=r)
    S7d(xg7, yv)


g = ab.solve(op7, n7, 4, 16, "LA", tol=1e-5, mxiter=3000, resid=r7, comm=comm, dtype=np.float32)
o = oracle().solve(lambda x: (S7_loc @ host_allgather(x, cnts7)).astype(np.float32), n7, 4, 16, "LA", tol=1e-5,
                   mxiter=3000, resid=r7, c_abi_tol=True, dtype=np.float32)
rel = np.abs(g.d - o.d).max() / np.abs(o.d).max()
report("pssaupd_c laplace2d (FP32)", g.info == 0 and g.ierr == 0 and g.nconv == o.nconv and rel <= 1e-4,
       f"counts gpu={counts_of(g)} oracle={counts_of(o)} rel eig diff={rel:.1e}")

# ---- 8. a sweep cut short on every rank: start vector inside a 3-dimensional invariant subspace of diag(1..N) ----
# rnorm collapses in mid-sweep, the gated start of the next step trips the stop flag, the remaining kernels (and their
# fused reductions) of the batch are no-ops on all ranks, and pdsaitr's restart (pdgetv0, per-rank seeds) takes over
r8 = np.zeros(N)
r8[[3, N // 2 + 1, N - 7]] = [1.0, -2.0, 0.5]
r8 = r8[first:first + cnt]
L.ab200_reset_seed()
g = ab.solve(lambda x, y, *_: torch.mul(diag, x, out=y), cnt, 4, 12, "LM", tol=1e-10, mxiter=2000, resid=r8, comm=comm)
o = oracle().solve(lambda x: diag_h * x, cnt, 4, 12, "LM", tol=1e-10, mxiter=2000, resid=r8, c_abi_tol=True)
rel = np.abs(np.sort(g.d) - np.sort(o.d)).max() / np.abs(o.d).max()
report("pdsaupd_c breakdown + restart in mid-sweep", g.info == o.info and g.nconv == o.nconv and rel <= 1e-8 and
       o.stats["nrstrt"] > 0, f"info={g.info}/{o.info} nconv={g.nconv}/{o.nconv} oracle restarts of the factorisation="
       f"{o.stats['nrstrt']} rel eig diff={rel:.1e}")

st = ab.launch_stats()
print(f"[rank {rank}] launches={st} reductions over: {path}", flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
L.ab200_comm_destroy(comm)
dist.destroy_process_group()
if rank == 0:
    print(f"MULTIGPU_CHECK world={world} path={path}", "PASS" if int(flag.item()) == 1 else "FAIL", flush=True)
sys.exit(0 if int(flag.item()) == 1 else 1)
