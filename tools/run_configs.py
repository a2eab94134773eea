"""Time-to-solution runs of BASELINE.json's configs through the C-ABI (device-resident RCI loop), one JSON line each.

  python tools/run_configs.py 2|3|4|5 [--scale f] [--tol t] [--mxiter m]
  torchrun --nproc-per-node N tools/run_configs.py 3           (config 3, z-slab partition, pdsaupd_c + NCCL halos)

config 2: 2-D Laplacian 4096^2, dsaupd nev=10 ncv=40 'LA'       (converges very slowly by nature: budgeted window)
config 3: 3-D Laplacian 512^3, pdsaupd-style nev=20 ncv=64 'LA'
config 4: 2-D convection-diffusion nx=2048 rho=100, dnaupd nev=6 ncv=30 'LR'
config 5: SVD via dsaupd on A^T A, A random sparse 20M x 5M with 16 nnz/row (splitmix64 rule of SURVEY 8d), nev=16 ncv=48
          'LM'; A row-sharded over the ranks (all-gather x, local A and A^T, reduce-scatter), L2-sized shards per GPU
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arpack_ng_b200 as ab  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", type=int)
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the grid edge by this factor (1 = full size)")
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--mxiter", type=int, default=3000)
    ap.add_argument("--which", default=None)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    comm = None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = ab.nccl_comm_from_torch_distributed()
    sym = True
    extra = {}
    if args.config == 2:
        nx = int(4096 * args.scale)
        A = ab.CsrOperator.laplace2d(nx, nx)
        n, nev, ncv, which, op = A.n, 10, 40, args.which or "LA", A
        r0 = ab.hashed_start_vector(n)
        name = f"2-D Laplacian {nx}^2"
    elif args.config == 3:
        e = int(512 * args.scale)
        z0, nzloc = ab.slab_partition(e, world, rank)
        A = ab.CsrOperator.laplace3d(e, e, e, z0=z0, nzloc=nzloc)
        n, nev, ncv, which = A.n, 20, 64, args.which or "LA"
        op = A  # under a communicator solve() applies it with the halo exchange
        r0 = ab.hashed_start_vector(n, i0=z0 * e * e)
        name = f"3-D Laplacian {e}^3 over {world} GPU(s)"
    elif args.config == 4:
        nx = int(2048 * args.scale)
        A = ab.CsrOperator.convdiff2d(nx, 100.0)
        n, nev, ncv, which, op, sym = A.n, 6, 30, args.which or "LR", A, False
        r0 = ab.hashed_start_vector(n)
        name = f"2-D convection-diffusion {nx}^2 rho=100"
    elif args.config == 5:
        m, k, per = int(20_000_000 * args.scale), int(5_000_000 * args.scale), 16
        k -= k % world
        # A row-sharded over the ranks (and cut into L2-sized shards on each), vectors sharded k/world per rank
        op = ab.GramOperator.randsparse(m, k, per, comm=comm)
        n, nev, ncv, which = op.n, 16, 48, args.which or "LM"
        r0 = ab.hashed_start_vector(n, i0=rank * n)
        name = f"SVD A^T A, A random sparse {m}x{k} (splitmix64 rule), {per} nnz/row, row-sharded over {world} GPU(s)"
        extra["nnz"] = m * per
        extra["shards_on_this_rank"] = len(op._keep)
    else:
        raise SystemExit("config must be 2..5")
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    st0 = ab.launch_stats()
    ab.profile(enable=True, reset=True)
    t0 = time.perf_counter()
    res = ab.solve(op, n, nev, ncv, which, sym=sym, tol=args.tol, mxiter=args.mxiter, resid=r0, comm=comm)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    dt = time.perf_counter() - t0
    prof = ab.profile(enable=False)
    st1 = ab.launch_stats()
    out = {"config": args.config, "problem": name, "n_local": n, "nev": nev, "ncv": ncv, "which": which, "tol": args.tol,
           "n_gpus": world, "info": res.info, "ierr": res.get("ierr"), "nconv": res.nconv,
           "restarts": int(res.iparam[2]), "nopx": int(res.iparam[8]), "nrorth": int(res.iparam[10]),
           "time_to_solution_s": dt, "lanczos_steps_per_s": int(res.iparam[8]) / dt,
           "gpu_launches": st1["kernels"] - st0["kernels"], "allreduces": st1["allreduces"] - st0["allreduces"],
           "kernels": {k: {"launches": v["launches"], "ms": round(v["ms"], 2),
                           "GBps": round(v["bytes"] / v["ms"] / 1e6, 1) if v["ms"] > 0 else None}
                       for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}, **extra}
    if sym and "d" in res:
        out["eigenvalues"] = [float(x) for x in res.d]
        if args.config in (2, 3, 4) and world == 1:
            rn = A.residuals(res.d, res.z, n)
            out["max_residual"] = float(rn.max())
        if args.config == 5:
            out["singular_values"] = [float(np.sqrt(max(x, 0.0))) for x in res.d]
    elif not sym and "dr" in res:
        out["eigenvalues_re"] = [float(x) for x in res.dr[:res.nconv]]
        out["eigenvalues_im"] = [float(x) for x in res.di[:res.nconv]]
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
