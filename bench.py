#!/usr/bin/env python
"""bench.py -- Lanczos steps/s of the dsaupd hot path on BASELINE.json's config 2.

  python bench.py --gpus N --steps K --warmup W          our arm (CUDA, one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference ...                   the reference algorithm on the host cores (oracle port)

One bench "step" = one dsaupd_c run with a fixed restart budget (--restarts R: nev + R*(ncv - kev) Lanczos steps,
exits with info = 1) on the 2-D 5-point Laplacian nx x nx (n = nx^2, CSR FP64), nev=10, ncv=40, which='LA',
tol=1e-10, start vector = the hashed vector of SURVEY.md §8(d).  value = OP*x count (iparam(9)) / time.

  value : device-resident path -- A, resid, V, workd live in HBM, the ido=+-1 hand-off passes device pointers to
          the CSR SpMV kernel; timed with CUDA events on the library's stream, max over ranks.
  e2e   : the same solve through dsaupd_c with everything the caller owns in HOST (pinned) memory -- the CSR matrix,
          resid, V, workd.  The caller registers its host matrix (ab200_register_csr_op_f64, one call before ido = 0);
          A and resid are uploaded, the solve runs in one dsaupd_c call, V and resid come back at ido = 99, all inside
          the timed region.  e2e_rci_handoff is the same with the unmodified reverse-communication loop: every hand-off
          crosses PCIe (library: D2H x, H2D y; OP: H2D x, SpMV, D2H y).  Byte counts are those of the copies issued.
  roofline : the dominant kernel of the timed region (by accumulated CUDA-event time), achieved = algorithmic bytes
          charged per launch / event time, against MEASURED_PEAKS.json's hbm_gbs.
  cpu_baseline : the oracle port (oracle/libref_arpack.so + OpenBLAS, all host threads) on a bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

NEV, NCV, WHICH, TOL = 10, 40, "LA", 1e-10


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_solve_sample(nx, restarts, threads, steps=1, warmup=0):
    """dsaupd (oracle restatement of SRC/dsaupd.f..dsapps.f) + threaded CSR SpMV on nx x nx, fixed restart budget."""
    from backends import lib as oracle_lib
    import arpack_ng_b200 as ab
    L = oracle_lib()
    n = nx * nx
    nnz = 5 * n - 4 * nx
    rowptr = np.empty(n + 1, dtype=np.int32)
    col = np.empty(nnz, dtype=np.int32)
    val = np.empty(nnz, dtype=np.float64)
    ip, dp = (lambda a: a.ctypes.data_as(C.POINTER(C.c_int))), (lambda a: a.ctypes.data_as(C.POINTER(C.c_double)))
    assert L.ref_gen_laplace2d(nx, nx, 1.0, ip(rowptr), ip(col), dp(val)) == nnz
    L.ref_set_blas_threads(threads)
    r0 = ab.hashed_start_vector_numpy(n)
    v = np.zeros(n * NCV)
    counts = np.zeros(5, dtype=np.int32)
    times = []
    nopx = 0
    for it in range(warmup + steps):
        ctx = C.c_void_p(L.ref_ctx_new())
        resid = r0.copy()
        tt, top = C.c_double(), C.c_double()
        info = L.ref_dsaupd_csr_solve(ctx, n, ip(rowptr), ip(col), dp(val), WHICH.encode(), NEV, NCV, TOL, restarts, 1,
                                      dp(resid), dp(v), None, None, None, ip(counts), threads, C.byref(tt),
                                      C.byref(top))
        L.ref_ctx_free(ctx)
        if it >= warmup:
            times.append(tt.value)
            nopx = int(counts[2])
    return {"seconds": times, "nopx": nopx, "info": int(info), "restarts": int(counts[0]), "nrorth": int(counts[4])}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    restarts = 1  # bounded sample: nev + one restart sweep = 40 OP*x on the full-size operator
    r = cpu_solve_sample(args.nx, restarts, threads, steps=args.steps, warmup=min(args.warmup, 1))
    total = sum(r["seconds"])
    value = r["nopx"] * len(r["seconds"]) / total
    sample = (f"same operator (2-D Laplacian {args.nx}x{args.nx}, CSR FP64) and solver parameters, restart budget "
              f"mxiter={restarts} ({r['nopx']} OP*x per step) instead of {args.restarts}")
    out = {"impl": "reference", "metric": "lanczos_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / len(r["seconds"]),
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, restarts),
           "cpu_baseline": {"value": value, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "reference = arpack-ng's dsaupd algorithm restated in C (oracle/) on OpenBLAS, all host threads; the "
                   "Fortran reference itself cannot be compiled in this image (no Fortran compiler)"}
    print(json.dumps(out))


def workload_config(args, restarts):
    if getattr(args, "workload", "laplace2d") == "laplace3d":
        e = args.nx
        return {"workload": f"BASELINE config 3: pdsaupd-style solve on 3-D 7-point Laplacian {e}^3 (n={e ** 3}) CSR FP64, "
                            f"z-slab row partition, nev=20 ncv=64 which=LA tol={TOL}, fixed restart budget",
                "nx": e, "n": e ** 3, "nev": 20, "ncv": 64, "which": "LA", "tol": TOL, "restarts_per_step": restarts,
                "start_vector": "splitmix64 hash, info=1",
                "l2": "inputs larger than L2 (V alone is %.1f GB)" % (e ** 3 * 64 * 8 / 1e9)}
    return {"workload": f"BASELINE config 2: dsaupd on 2-D 5-point Laplacian {args.nx}x{args.nx} (n={args.nx * args.nx}) "
                        f"CSR FP64, nev={NEV} ncv={NCV} which={WHICH} tol={TOL}, fixed restart budget",
            "nx": args.nx, "n": args.nx * args.nx, "nev": NEV, "ncv": NCV, "which": WHICH, "tol": TOL,
            "restarts_per_step": restarts, "start_vector": "splitmix64 hash, info=1",
            "l2": "inputs larger than L2 (V alone is %.1f GB)" % (args.nx * args.nx * NCV * 8 / 1e9)}


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import arpack_ng_b200 as ab
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = ab.nccl_comm_from_torch_distributed()
    L = ab.lib()
    nx = args.nx
    nev, ncv = NEV, NCV
    # block-row partition, PARPACK's layout (dsaupd.f:331-349): y-slabs of the 2-D grid / z-slabs of the 3-D grid
    y0, nyloc = ab.slab_partition(nx, world, rank)
    if args.workload == "laplace3d":
        nev, ncv = 20, 64
        A = ab.CsrOperator.laplace3d(nx, nx, nx, z0=y0, nzloc=nyloc)
        r0 = ab.hashed_start_vector(A.n, i0=y0 * nx * nx)
    elif world == 1:
        A = ab.CsrOperator.laplace2d(nx, nx)
        r0 = ab.hashed_start_vector(A.n)
    else:
        A = ab.CsrOperator.laplace3d(nx, 1, nx, z0=y0, nzloc=nyloc, diag=4.0)
        r0 = ab.hashed_start_vector(A.n, i0=y0 * nx)
    op = A  # world > 1: solve() applies it with the halo exchange (CsrOperator.apply_halo_ptr) on the library's comm
    n = A.n
    restarts = args.restarts

    registered = args.op_mode == "registered"

    host_arrays = [None]
    # the caller's V/workd/resid are allocated once, outside the timed region, like the arrays a reference driver
    # declares (EXAMPLES/SIMPLE/dssimp.f:213-219); every solve of the bench reuses them
    dev_arrays = ab.alloc_device_buffers(n, ncv)

    def one_solve(host_buffers=False, resid=None, reg=None):
        reg = registered if reg is None else reg
        if host_buffers and host_arrays[0] is None:
            host_arrays[0] = ab.alloc_host_buffers(n, ncv)
        return ab.solve(op, n, nev, ncv, WHICH, tol=TOL, mxiter=restarts, resid=resid if resid is not None else r0,
                        eupd=False, host_buffers=host_buffers, comm=comm,
                        buffers=host_arrays[0] if host_buffers else dev_arrays,
                        registered_op=A if (reg and not host_buffers) else None)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        res = one_solve()
    barrier()
    ab.profile(enable=True, reset=True)
    st0 = ab.launch_stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    nopx = 0
    for _ in range(args.steps):
        res = one_solve()
        nopx += int(res.iparam[8])
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    elapsed = ev0.elapsed_time(ev1) / 1e3
    prof = ab.profile(enable=False)
    st1 = ab.launch_stats()
    if dist is not None:
        t = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    value = nopx / elapsed

    # ---- the opt-in registered-operator mode (one *aupd_c call per solve, K1+K2+K3 fused), reported beside the
    # strict-RCI headline; same operator, same restart budget, device-resident
    reg_mode = None
    # (N = 1 only; for N > 1 run `bench.py --op-mode registered`, which times the registered mode in the main region)
    if world == 1 and not registered and not args.no_registered:
        for _ in range(2):
            one_solve(reg=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nopr = 0
        for _ in range(args.steps):
            rr = one_solve(reg=True)
            nopr += int(rr.iparam[8])
        e1.record()
        barrier()
        reg_s = e0.elapsed_time(e1) / 1e3
        if dist is not None:
            t = torch.tensor([reg_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            reg_s = float(t.item())
        reg_mode = {"value": nopr / reg_s, "unit": "steps/s",
                    "ms_per_lanczos_step": 1e3 * reg_s / nopr, "aupd_calls_per_solve": 1,
                    "fused_dot_maxdiff": rr.fused_dot_maxdiff,
                    "note": "ab200_register_csr_op_f64 (ab200_register_csr_halo_op_f64 under a communicator): OP applied "
                            "inside *aupd_c, v_j scaling and alpha/||w||^2 fused into the SpMV kernel"}

    # ---- e2e: HOST buffers through the reference-facing C-ABI (N = 1 only: one PCIe link per GPU anyway) ----
    # Everything the caller owns starts and ends in (pinned) host memory, every copy is inside the timed region.
    #   e2e              : the caller also owns the CSR matrix on the host and registers it (one extra call before
    #                      ido = 0, ab200_register_csr_op_f64 with host arrays): upload of A + resid, the whole solve in
    #                      one dsaupd_c call, download of V + resid.
    #   e2e_rci_handoff  : the unmodified reverse-communication loop -- every ido = 1 hand-off crosses PCIe
    #                      (library: D2H x, H2D y; the caller's GPU OP: H2D x, SpMV, D2H y).
    e2e = None
    e2e_rci = None
    if world == 1 and not args.no_e2e:
        r0h = r0.cpu().numpy()
        e2e_steps = max(1, min(args.steps, 2))
        w = 8

        def host_arm(registered_host):
            kw = {}
            if registered_host:
                kw = dict(registered_op=host_csr[0])
            if host_arrays[0] is None:
                host_arrays[0] = ab.alloc_host_buffers(n, ncv)

            def run():
                return ab.solve(None if registered_host else op, n, nev, ncv, WHICH, tol=TOL, mxiter=restarts, resid=r0h,
                                eupd=False, host_buffers=True, buffers=host_arrays[0], **kw)
            run()  # warm-up (pinned allocation, page faults)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            cnt = 0
            for _ in range(e2e_steps):
                rr = run()
                cnt += int(rr.iparam[8])
                if registered_host and (rr.nsteps != 0 or int(rr.iparam[8]) != nopx // args.steps):
                    raise RuntimeError("registered host-CSR solve took a different path than the device-resident one")
                if not registered_host and rr.nsteps != int(rr.iparam[8]):
                    raise RuntimeError("e2e solve did not hand every OP*x to the caller")
            torch.cuda.synchronize()
            return cnt, time.perf_counter() - t0

        per = nopx // args.steps
        host_csr = [None]
        try:
            host_csr[0] = ab.HostCsr.from_operator(A)   # the caller's matrix, in pinned host memory, outside the timing
            cnt, dt = host_arm(True)
            # library: H2D rowptr/col/val + resid at ido = 0; D2H V + resid at ido = 99; nothing per Lanczos step
            e2e = {"value": cnt / dt, "unit": "steps/s", "h2d_bytes_per_step": int(host_csr[0].nbytes() + n * w),
                   "d2h_bytes_per_step": int(n * ncv * w + n * w), "bench_steps": e2e_steps,
                   "aupd_calls_per_solve": 1,
                   "buffers": "pinned host CSR arrays + resid/V/workd; ab200_register_csr_op_f64(host arrays) then one "
                              "dsaupd_c call: A and resid uploaded, V and resid downloaded inside the timed region"}
        except Exception as ex:  # keep the bench line alive: the hand-off arm below then is the e2e number
            e2e = None
            e2e_err = repr(ex)
        host_csr[0] = None
        cnt, dt = host_arm(False)
        # library: H2D resid once; per hand-off D2H x + H2D y; at ido=99 D2H V + resid.  OP: H2D x + D2H y per call
        h2d = n * w + per * (n * w) + per * (n * w)
        d2h = per * (n * w) + per * (n * w) + n * ncv * w + n * w
        e2e_rci = {"value": cnt / dt, "unit": "steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "bench_steps": e2e_steps, "buffers": "pinned host resid/V/workd, unmodified RCI loop, "
                                                        "OP = H2D + CSR SpMV kernel + D2H"}
        if e2e is None:
            e2e = dict(e2e_rci, registered_arm_error=e2e_err)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    if dist is not None:
        dist.barrier()
    peak, peak_src = peaks()
    # dominant kernel by accumulated event time
    roof = None
    if prof:
        name, top = max(prof.items(), key=lambda kv: kv[1]["ms"])
        total_ms = sum(v["ms"] for v in prof.values())
        ach = top["bytes"] / (top["ms"] * 1e-3) / 1e9 if top["ms"] > 0 else 0.0
        traffic = None
        traffic_ref = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            try:
                ent = json.load(open(tp)).get(name, {})
                traffic = ent.get("dram_bytes_per_launch")
                # the ncu capture is one launch at a fixed column count j; its algorithmic bytes are recorded with it
                traffic_ref = {"algorithmic_bytes_at_capture": ent.get("algorithmic_bytes_at_capture"),
                               "traffic_over_algorithmic": ent.get("traffic_over_algorithmic")}
            except Exception:
                traffic = None
        step_bytes = sum(v["bytes"] for v in prof.values())
        roof = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_ref": traffic_ref, "peak_source": peak_src,
                "peak_note": "the denominator is a measured COPY bandwidth (equal read and write streams); kernels that "
                             "mostly read (multi-dots, updates of one vector against j columns) can exceed it", "launches": top["launches"],
                "avg_launch_ms": top["ms"] / max(1, top["launches"]), "share_of_kernel_time": top["ms"] / total_ms,
                "algorithmic_bytes_per_launch": top["bytes"] / max(1, top["launches"]),
                "all_kernels": {k: {"launches": v["launches"], "ms": round(v["ms"], 3),
                                    "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None}
                                for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])},
                "lanczos_step_aggregate": {"algorithmic_GB_per_s_per_gpu": step_bytes / elapsed / 1e9,
                                           "frac_of_peak": step_bytes / elapsed / 1e9 / peak,
                                           "kernel_time_share_of_elapsed": total_ms * 1e-3 / elapsed}}
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        c = cpu_solve_sample(nx, 1, threads)
        cv = c["nopx"] / c["seconds"][0]
        cpu = {"value": cv, "unit": "steps/s", "cores": threads, "kind": "port",
               "sample": f"same operator and parameters with restart budget mxiter=1 ({c['nopx']} OP*x, "
                         f"{c['seconds'][0]:.1f} s), oracle port + OpenBLAS, {threads} threads"}
    out = {"metric": "lanczos_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, restarts), "lanczos_steps_per_bench_step": nopx // args.steps,
           "ms_per_lanczos_step": 1e3 * elapsed / nopx, "info": int(res.info), "wall_s": wall,
           "gpu_launches": st1["kernels"] - st0["kernels"], "allreduces": st1["allreduces"] - st0["allreduces"],
           "kernel_path": {"tma": st1["tma_path"] - st0["tma_path"], "generic": st1["generic_path"] - st0["generic_path"]},
           "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "e2e_rci_handoff": e2e_rci, "clocks": clocks,
           "op_mode": "registered" if registered else "rci", "registered_op_mode": reg_mode,
           "allreduce_path": (None if comm is None else ("peer-memory kernel" if L.ab200_comm_uses_p2p(comm) else "nccl"))}
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=4096)
    ap.add_argument("--restarts", type=int, default=20,
                    help="restart budget of one bench step (SURVEY.md 8d: a fixed 20-restart window)")
    ap.add_argument("--workload", default="laplace2d", choices=["laplace2d", "laplace3d"],
                    help="laplace2d = BASELINE config 2 (default, the headline); laplace3d = config 3 (use --nx 512)")
    ap.add_argument("--op-mode", default="rci", choices=["rci", "registered"],
                    help="rci = the reference's reverse-communication loop (headline); registered = opt-in "
                         "ab200_register_csr_op mode for the timed region (N=1)")
    ap.add_argument("--no-registered", action="store_true", help="skip the extra registered-operator measurement")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
