#!/usr/bin/env python
"""bench.py -- Lanczos steps/s of the dsaupd hot path on BASELINE.json's config 2.

  python bench.py --gpus N --steps K --warmup W          our arm (CUDA, one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference ...                   the reference algorithm on the host cores (oracle port)

One bench "step" = one dsaupd_c run with a fixed restart budget (--restarts R: nev + R*(ncv - kev) Lanczos steps,
exits with info = 1) on the 2-D 5-point Laplacian nx x nx (n = nx^2, CSR FP64), nev=10, ncv=40, which='LA',
tol=1e-10, start vector = the hashed vector of SURVEY.md §8(d).  value = OP*x count (iparam(9)) / time.

  value : device-resident path -- A, resid, V, workd live in HBM, the ido=+-1 hand-off passes device pointers to
          the CSR SpMV kernel (the library enqueues a whole sweep without waiting for the GPU); timed with CUDA events
          on the library's stream, max over ranks, WITHOUT the per-kernel profiler.
  e2e   : the SAME reverse-communication loop with the caller's resid, V, workd in HOST (pinned) memory: every hand-off
          crosses PCIe (library: D2H x, H2D y; the caller's GPU OP: H2D x, SpMV, D2H y), V and resid come back at
          ido = 99 -- all inside the timed region.  e2e_registered_host_csr: the caller also owns the CSR matrix on the
          host and registers it (ab200_register_csr_op_f64): A and resid up, one dsaupd_c call, V and resid down.
  value_mxiter1 : the GPU on the restart budget of the reference arm (mxiter = 1).
  roofline : the dominant kernel by accumulated CUDA-event time of ONE extra solve after the timed region (same
          launches), achieved = algorithmic bytes charged per launch / event time, against MEASURED_PEAKS.json's hbm_gbs.
  cpu_baseline : the oracle port (oracle/libref_arpack.so + OpenBLAS, all host threads) on a bounded sample;
  fullsize_parity : that run's projected matrix, Ritz values, bounds and counts against a GPU run of the same budget.
  config3 : north_star's target (3-D 7-point Laplacian 512^3, nev 20, ncv 64, z-slabs over the ranks): steps/s on a
          fixed restart budget, per-step HBM fraction, kernel share; time to solution from 4 GPUs up.
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
def cpu_solve_sample(nx, restarts, threads, steps=1, warmup=0, want_workl=False):
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
    workl = np.zeros(NCV * NCV + 8 * NCV)
    times = []
    nopx = 0
    for it in range(warmup + steps):
        ctx = C.c_void_p(L.ref_ctx_new())
        resid = r0.copy()
        tt, top = C.c_double(), C.c_double()
        info = L.ref_dsaupd_csr_solve(ctx, n, ip(rowptr), ip(col), dp(val), WHICH.encode(), NEV, NCV, TOL, restarts, 1,
                                      dp(resid), dp(v), None, None, dp(workl) if want_workl else None, ip(counts),
                                      threads, C.byref(tt), C.byref(top))
        L.ref_ctx_free(ctx)
        if it >= warmup:
            times.append(tt.value)
            nopx = int(counts[2])
    return {"seconds": times, "nopx": nopx, "info": int(info), "restarts": int(counts[0]), "nconv": int(counts[1]),
            "nrorth": int(counts[4]), "workl": workl}


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
def _hbm_bytes(prof):
    return sum(v["bytes"] for v in prof.values())


def _kernels(prof):
    """the profiled launches without the profiler's pseudo-entries such as '(between launches)'."""
    return {k: v for k, v in prof.items() if not k.startswith("(")}


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
    peak, peak_src = peaks()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def make_workload(kind, nx):
        """(operator, start vector, nev, ncv): block-row partition, PARPACK's layout (dsaupd.f:331-349)."""
        y0, nyloc = ab.slab_partition(nx, world, rank)
        if kind == "laplace3d":
            A = ab.CsrOperator.laplace3d(nx, nx, nx, z0=y0, nzloc=nyloc)
            return A, ab.hashed_start_vector(A.n, i0=y0 * nx * nx), 20, 64
        if world == 1:
            A = ab.CsrOperator.laplace2d(nx, nx)
            return A, ab.hashed_start_vector(A.n), NEV, NCV
        A = ab.CsrOperator.laplace3d(nx, 1, nx, z0=y0, nzloc=nyloc, diag=4.0)   # y-slabs of the 2-D grid
        return A, ab.hashed_start_vector(A.n, i0=y0 * nx), NEV, NCV

    def timed(fn, steps, warmup):
        """`steps` calls of fn() between barriers, CUDA events on the library's stream, max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        nopx, res = 0, None
        for _ in range(steps):
            res = fn()
            nopx += int(res.iparam[8])
        ev1.record()
        barrier()
        wall = time.perf_counter() - t0
        return nopx, max_over_ranks(ev0.elapsed_time(ev1) / 1e3), wall, res

    def profiled(fn):
        """One extra, untimed call with the per-kernel CUDA-event profiler on (two event records per launch perturb
        the host side, so this never runs inside a timed region)."""
        barrier()
        ab.profile(enable=True, reset=True)
        fn()
        torch.cuda.synchronize()
        prof = ab.profile(enable=False)
        barrier()
        return prof

    # ================================ config 2: the headline ================================
    A, r0, nev, ncv = make_workload(args.workload, args.nx)
    n = A.n
    restarts = args.restarts
    registered = args.op_mode == "registered"
    dev_arrays = ab.alloc_device_buffers(n, ncv)   # the caller's V/workd/resid (dssimp.f:213-219), reused by every solve
    host_arrays = [None]

    def one_solve(reg=None, mx=None, eupd=False):
        reg = registered if reg is None else reg
        return ab.solve(A, n, nev, ncv, WHICH, tol=TOL, mxiter=mx or restarts, resid=r0, eupd=eupd, comm=comm,
                        buffers=dev_arrays, registered_op=A if reg else None)

    sampler = ClockSampler(local)
    st0 = ab.launch_stats()
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        one_solve()
    st0 = ab.launch_stats()
    rt0 = ab.host_round_trips()
    if args.profile_in_timed:   # diagnostic only: the r1 way (events around every launch inside the timed region)
        ab.profile(enable=True, reset=True)
    nopx, elapsed, wall, res = timed(one_solve, args.steps, 0)
    prof_timed = ab.profile(enable=False) if args.profile_in_timed else None
    st1 = ab.launch_stats()
    rt1 = ab.host_round_trips()
    clocks = sampler.stop() if rank == 0 else None
    value = nopx / elapsed
    prof = profiled(one_solve)

    # a GPU figure on the reference arm's own budget (mxiter = 1), so that the two arms can be compared like for like
    mx1 = None
    if not args.no_extras:
        n1, t1, _, r1 = timed(lambda: one_solve(mx=1), max(2, min(args.steps, 5)), 1)
        mx1 = {"value": n1 / t1, "unit": "steps/s", "restarts_per_step": 1,
               "lanczos_steps_per_bench_step": int(r1.iparam[8]),
               "note": "same solve with the restart budget of the --impl reference arm (mxiter=1)"}

    # ---- the opt-in registered-operator mode (one *aupd_c call per solve, K1+K2+K3 fused), same step count
    reg_mode = None
    if not registered and not args.no_registered:
        nr, tr, _, rr = timed(lambda: one_solve(reg=True), args.steps, 2)
        reg_mode = {"value": nr / tr, "unit": "steps/s", "ms_per_lanczos_step": 1e3 * tr / nr,
                    "aupd_calls_per_solve": 1, "bench_steps": args.steps, "fused_dot_maxdiff": rr.fused_dot_maxdiff,
                    "note": "ab200_register_csr_op_f64 (ab200_register_csr_halo_op_f64 under a communicator): OP applied "
                            "inside *aupd_c, v_j scaling and alpha/||w||^2 fused into the SpMV kernel"}

    # ---- e2e: HOST buffers through the reference-facing C-ABI (N = 1: one PCIe link per GPU anyway) ----
    #   e2e                     : the SAME code path as `value` -- the unmodified reverse-communication loop -- with the
    #                             caller's resid/V/workd in pinned host memory: every ido = 1 hand-off crosses PCIe
    #                             (library: D2H x, H2D y; the caller's GPU OP: H2D x, SpMV, D2H y), V + resid come back at
    #                             ido = 99.  PCIe-bound by the protocol, not by the kernels.
    #   e2e_registered_host_csr : the caller also owns the CSR matrix on the host and registers it (one extra call,
    #                             ab200_register_csr_op_f64 with host arrays): upload of A + resid, the whole solve in one
    #                             dsaupd_c call, download of V + resid -- all inside the timed region.
    e2e = None
    e2e_reg = None
    if world == 1 and not args.no_e2e:
        r0h = r0.cpu().numpy()
        w = 8
        per = nopx // args.steps
        host_arrays[0] = ab.alloc_host_buffers(n, ncv)

        def host_solve(host_csr=None):
            kw = dict(registered_op=host_csr) if host_csr is not None else {}
            return ab.solve(None if host_csr is not None else A, n, nev, ncv, WHICH, tol=TOL, mxiter=restarts, resid=r0h,
                            eupd=False, host_buffers=True, buffers=host_arrays[0], **kw)

        def wall_timed(fn, steps):
            fn()  # warm-up (pinned allocation, page faults)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            cnt, rr = 0, None
            for _ in range(steps):
                rr = fn()
                cnt += int(rr.iparam[8])
            torch.cuda.synchronize()
            return cnt, time.perf_counter() - t0, rr

        # the hand-off loop moves 4 n doubles per Lanczos step at ~55 GB/s: keep the leg inside ~170 s
        est = per * (4 * n * w / 50e9 + elapsed / nopx)
        e2e_steps = max(1, min(args.steps, int(170.0 / max(est, 1e-3))))
        cnt, dt, rr = wall_timed(host_solve, e2e_steps)
        if rr.nsteps != int(rr.iparam[8]) or int(rr.iparam[8]) != per:
            raise RuntimeError("e2e solve took a different path than the device-resident one")
        # library: H2D resid once; per hand-off D2H x + H2D y; at ido=99 D2H V + resid.  OP: H2D x + D2H y per call
        e2e = {"value": cnt / dt, "unit": "steps/s", "h2d_bytes_per_step": int(n * w + 2 * per * n * w),
               "d2h_bytes_per_step": int(2 * per * n * w + n * ncv * w + n * w), "bench_steps": e2e_steps,
               "path": "strict RCI (same as value)",
               "buffers": "pinned host resid/V/workd, unmodified reverse-communication loop; the caller's OP = H2D x + "
                          "CSR SpMV kernel + D2H y"}
        try:
            hc = ab.HostCsr.from_operator(A)   # the caller's matrix, in pinned host memory, outside the timing
            cnt, dt, rr = wall_timed(lambda: host_solve(hc), args.steps)
            if rr.nsteps != 0 or int(rr.iparam[8]) != per:
                raise RuntimeError("registered host-CSR solve took a different path than the device-resident one")
            e2e_reg = {"value": cnt / dt, "unit": "steps/s", "h2d_bytes_per_step": int(hc.nbytes() + n * w),
                       "d2h_bytes_per_step": int(n * ncv * w + n * w), "bench_steps": args.steps,
                       "aupd_calls_per_solve": 1,
                       "buffers": "pinned host CSR arrays + resid/V/workd; ab200_register_csr_op_f64(host arrays) then "
                                  "one dsaupd_c call: A and resid uploaded, V and resid downloaded inside the timed region"}
            del hc
        except Exception as ex:  # keep the bench line alive
            e2e_reg = {"error": repr(ex)}
        host_arrays[0] = None

    # ---- cpu_baseline + full-size parity on a fixed budget (rank 0, N = 1): the oracle runs the full-size operator for
    # mxiter = 1 anyway; its projected matrix, Ritz values, bounds and counts are compared with a GPU run of the same
    # budget
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu and args.workload == "laplace2d":
        threads = os.cpu_count() or 1
        c = cpu_solve_sample(args.nx, 1, threads, want_workl=True)
        cv = c["nopx"] / c["seconds"][0]
        cpu = {"value": cv, "unit": "steps/s", "cores": threads, "kind": "port",
               "sample": f"same operator and parameters with restart budget mxiter=1 ({c['nopx']} OP*x, "
                         f"{c['seconds'][0]:.1f} s), oracle port + OpenBLAS, {threads} threads"}
        g = one_solve(mx=1)
        wo, wg = c["workl"], g.workl
        ih, ir, ib = int(g.ipntr[4]) - 1, int(g.ipntr[5]) - 1, int(g.ipntr[6]) - 1

        def rel(a, b):
            return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
        parity = {"budget": "mxiter=1, tol=%g" % TOL, "info": [int(g.info), c["info"]],
                  "counts_gpu": {"restarts": int(g.iparam[2]), "nconv": int(g.iparam[4]), "nopx": int(g.iparam[8]),
                                 "nrorth": int(g.iparam[10])},
                  "counts_oracle": {"restarts": c["restarts"], "nconv": c["nconv"], "nopx": c["nopx"],
                                    "nrorth": c["nrorth"]},
                  "H_max_rel_diff": rel(wg[ih:ih + 2 * ncv], wo[ih:ih + 2 * ncv]),
                  "ritz_max_rel_diff": rel(wg[ir:ir + ncv], wo[ir:ir + ncv]),
                  "bounds_max_abs_diff_over_norm": float(np.abs(wg[ib:ib + ncv] - wo[ib:ib + ncv]).max()
                                                        / max(np.abs(wo[ir:ir + ncv]).max(), 1e-300))}
        parity["counts_identical"] = parity["counts_gpu"] == parity["counts_oracle"]
        parity["ok"] = bool(parity["counts_identical"] and parity["info"][0] == parity["info"][1]
                            and parity["H_max_rel_diff"] < 1e-10 and parity["ritz_max_rel_diff"] < 1e-10)

    # ================================ config 3 (north_star's target) ================================
    cfg3 = None
    if not args.no_config3 and args.workload == "laplace2d":
        del dev_arrays, A, r0
        torch.cuda.empty_cache()
        e3 = args.nx3
        try:
            A3, r03, nev3, ncv3 = make_workload("laplace3d", e3)
            buf3 = ab.alloc_device_buffers(A3.n, ncv3)

            def solve3(eupd=False, mx=None):
                return ab.solve(A3, A3.n, nev3, ncv3, "LA", tol=TOL, mxiter=mx or args.restarts3, resid=r03, eupd=eupd,
                                comm=comm, buffers=buf3)
            n3, t3, _, r3 = timed(solve3, max(1, min(args.steps, 2)), 1)
            p3 = profiled(solve3)
            b3 = _hbm_bytes(p3)
            p3 = _kernels(p3)
            k3 = sum(v["ms"] for v in p3.values()) * 1e-3
            per3 = int(r3.iparam[8])
            cfg3 = {"workload": f"BASELINE config 3: pdsaupd-style solve on 3-D 7-point Laplacian {e3}^3 (n={e3 ** 3}) CSR "
                                f"FP64, z-slab row partition over {world} GPU(s), nev={nev3} ncv={ncv3} which=LA tol={TOL}, "
                                f"fixed budget of {args.restarts3} restarts",
                    "value": n3 / t3, "unit": "steps/s", "ms_per_lanczos_step": 1e3 * t3 / n3,
                    "lanczos_steps_per_bench_step": per3, "bench_steps": max(1, min(args.steps, 2)), "info": int(r3.info),
                    "step_hbm": {"algorithmic_GB_per_s_per_gpu": b3 / (t3 / max(1, min(args.steps, 2))) / 1e9,
                                 "frac_of_peak": b3 / (t3 / max(1, min(args.steps, 2))) / 1e9 / peak,
                                 "kernel_time_share_of_elapsed": k3 / (t3 / max(1, min(args.steps, 2)))},
                    "kernels": {k: {"launches": v["launches"], "ms": round(v["ms"], 2),
                                    "GBps": round(v["bytes"] / v["ms"] / 1e6, 1) if v["ms"] > 0 else None}
                                for k, v in sorted(p3.items(), key=lambda kv: -kv[1]["ms"])}}
            # time to solution of north_star's target: ~420 s on one GPU, ~53 s on eight -- run by default from 4 GPUs up
            if (args.tts3 or world >= 4) and not args.no_tts3:
                barrier()
                t0 = time.perf_counter()
                rs = solve3(eupd=True, mx=3000)
                torch.cuda.synchronize()
                barrier()
                cfg3["time_to_solution"] = {"seconds": max_over_ranks(time.perf_counter() - t0), "tol": TOL,
                                            "info": int(rs.info), "nconv": int(rs.nconv),
                                            "restarts": int(rs.iparam[2]), "nopx": int(rs.iparam[8])}
            del A3, r03, buf3
        except Exception as ex:
            cfg3 = {"error": repr(ex)}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    if dist is not None:
        dist.barrier()
    # dominant kernel by accumulated event time (profiled pass: same solve, same launches)
    roof = None
    if prof:
        gaps = prof.get("(between launches)")
        prof = _kernels(prof)
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
        step_bytes = _hbm_bytes(prof)            # algorithmic bytes of ONE solve
        per_solve = elapsed / args.steps         # un-profiled time of one solve
        roof = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_ref": traffic_ref, "peak_source": peak_src,
                "how": "CUDA events around every launch of one extra solve after the timed region (same restart budget, "
                       "same launches); the timed region itself runs without the profiler",
                "peak_note": "the denominator is a measured COPY bandwidth (equal read and write streams); kernels that "
                             "mostly read (multi-dots, updates of one vector against j columns) can exceed it",
                "launches": top["launches"],
                "avg_launch_ms": top["ms"] / max(1, top["launches"]), "share_of_kernel_time": top["ms"] / total_ms,
                "algorithmic_bytes_per_launch": top["bytes"] / max(1, top["launches"]),
                "all_kernels": {k: {"launches": v["launches"], "ms": round(v["ms"], 3),
                                    "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None}
                                for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])},
                "lanczos_step_aggregate": {"algorithmic_GB_per_s_per_gpu": step_bytes / per_solve / 1e9,
                                           "frac_of_peak": step_bytes / per_solve / 1e9 / peak,
                                           "kernel_time_share_of_elapsed": total_ms * 1e-3 / per_solve,
                                           "device_ms_between_launches_profiled_pass":
                                               (round(gaps["ms"], 3) if gaps else None)}}
    out = {"metric": "lanczos_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, restarts), "lanczos_steps_per_bench_step": nopx // args.steps,
           "ms_per_lanczos_step": 1e3 * elapsed / nopx, "info": int(res.info), "wall_s": wall,
           "gpu_launches": st1["kernels"] - st0["kernels"], "allreduces": st1["allreduces"] - st0["allreduces"],
           "host_round_trips_per_lanczos_step": (rt1 - rt0) / max(1, nopx),
           "kernel_path": {"tma": st1["tma_path"] - st0["tma_path"], "generic": st1["generic_path"] - st0["generic_path"]},
           "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "e2e_registered_host_csr": e2e_reg, "clocks": clocks,
           "op_mode": "registered" if registered else "rci", "registered_op_mode": reg_mode, "value_mxiter1": mx1,
           "fullsize_parity": parity, "config3": cfg3,
           "profile_in_timed": (None if not prof_timed else
                                {"kernel_ms_per_solve": sum(v["ms"] for v in _kernels(prof_timed).values()) / args.steps,
                                 "between_launches_ms_per_solve":
                                     prof_timed.get("(between launches)", {"ms": 0.0})["ms"] / args.steps,
                                 "kernel_time_share_of_elapsed":
                                     sum(v["ms"] for v in _kernels(prof_timed).values()) * 1e-3 / elapsed}),
           "allreduce_path": (None if comm is None else ("peer-memory" if L.ab200_comm_uses_p2p(comm) else "nccl"))}
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
    ap.add_argument("--no-extras", action="store_true", help="skip the mxiter=1 figure")
    ap.add_argument("--profile-in-timed", action="store_true", help="diagnostic: per-kernel events inside the timed region")
    ap.add_argument("--no-config3", action="store_true", help="skip the config-3 block (3-D Laplacian)")
    ap.add_argument("--nx3", type=int, default=512, help="grid edge of the config-3 block")
    ap.add_argument("--restarts3", type=int, default=3, help="restart budget of one config-3 solve")
    ap.add_argument("--tts3", action="store_true", help="also run config 3 to convergence (time to solution); default "
                                                        "from 4 GPUs up")
    ap.add_argument("--no-tts3", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
