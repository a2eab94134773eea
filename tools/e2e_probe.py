"""Where the time of a host-buffer (e2e) Lanczos step goes: library call (D2H x, H2D y, kernels) vs the caller's OP."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import arpack_ng_b200 as ab

nx = 4096
A = ab.CsrOperator.laplace2d(nx, nx)
n, nev, ncv = A.n, 10, 40
bufs = ab.alloc_host_buffers(n, ncv)
r0 = ab.hashed_start_vector(n).cpu().numpy()
t_op = [0.0, 0]


def op(x, y, *_):
    t0 = time.perf_counter()
    A(x, y)
    t_op[0] += time.perf_counter() - t0
    t_op[1] += 1


for it in range(2):
    t_op[:] = [0.0, 0]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = ab.solve(op, n, nev, ncv, "LA", tol=1e-10, mxiter=4, resid=r0, eupd=False, host_buffers=True, buffers=bufs)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    steps = int(r.iparam[8])
    print(json.dumps({"solve_s": dt, "steps": steps, "ms_per_step": 1e3 * dt / steps, "op_calls": t_op[1],
                      "op_ms_per_call": 1e3 * t_op[0] / t_op[1], "library_ms_per_step": 1e3 * (dt - t_op[0]) / steps,
                      "x_is_pinned": bool(bufs[1].is_pinned())}))
