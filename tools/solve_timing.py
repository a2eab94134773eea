"""Host-side timeline of one bench solve (config 2 by default): wall time of the whole ab.solve call, of the first
(ido = 0) call, of the release, of all *aupd_c re-entries and of all OP launches, next to the CUDA-event time; the
library prints its own phase clock (AB200_TIMING=1) on stderr.  Diagnostic for 'kernel share of elapsed' < 1."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

os.environ["AB200_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arpack_ng_b200 as ab  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nev, ncv, restarts = 10, 40, 20
L = ab.lib()
A = ab.CsrOperator.laplace2d(nx, nx)
n = A.n
r0 = ab.hashed_start_vector(n)
v, workd, res = ab.alloc_device_buffers(n, ncv)
c_int_p = C.POINTER(C.c_int)


def one():
    t = {}
    t0 = time.perf_counter()
    workl = np.zeros(ncv * ncv + 8 * ncv)
    iparam = np.zeros(11, dtype=np.int32)
    ipntr = np.zeros(14, dtype=np.int32)
    iparam[0], iparam[2], iparam[3], iparam[6] = 1, restarts, 1, 1
    ido = np.zeros(1, dtype=np.int32)
    info = np.ones(1, dtype=np.int32)
    res.copy_(r0)
    args = (ido.ctypes.data_as(c_int_p), b"I", n, b"LA", nev, C.c_double(1e-10), res.data_ptr(), ncv, v.data_ptr(), n,
            iparam.ctypes.data_as(c_int_p), ipntr.ctypes.data_as(c_int_p), workd.data_ptr(), workl.ctypes.data,
            int(workl.size), info.ctypes.data_as(c_int_p))
    wb = workd.data_ptr()
    t_aupd = t_op = 0.0
    t_first = None
    calls = 0
    while True:
        a = time.perf_counter()
        L.dsaupd_c(*args)
        b = time.perf_counter()
        if t_first is None:
            t_first = b - a
        else:
            t_aupd += b - a
        calls += 1
        if ido[0] in (1, -1):
            A.apply_ptr(wb + (int(ipntr[0]) - 1) * 8, wb + (int(ipntr[1]) - 1) * 8)
            t_op += time.perf_counter() - b
        else:
            break
    c = time.perf_counter()
    L.ab200_release(workl.ctypes.data)
    d = time.perf_counter()
    torch.cuda.synchronize()
    e = time.perf_counter()
    return {"total_s": e - t0, "first_call_s": t_first, "aupd_reentries_s": t_aupd, "op_launch_s": t_op,
            "release_s": d - c, "final_sync_s": e - d, "calls": calls, "nopx": int(iparam[8]),
            "loop_s": c - t0}


for _ in range(2):
    one()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
r = one()
e1.record()
torch.cuda.synchronize()
r["cuda_event_s"] = e0.elapsed_time(e1) / 1e3
print(json.dumps(r))
