"""PCIe probe for the host-buffer (e2e) path: pinned H2D / D2H bandwidth at the size of one hand-off vector
(n = 4096^2 doubles = 134 MB), and the time of one host-buffer Lanczos step split into its pieces."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import arpack_ng_b200 as ab

n = 4096 * 4096
h = torch.zeros(n, dtype=torch.float64).pin_memory()
d = torch.zeros(n, dtype=torch.float64, device="cuda")
out = {}
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    out[name + "_GBps"] = 10 * n * 8 / (time.perf_counter() - t0) / 1e9
# both directions at once (two streams)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.zeros(n, dtype=torch.float64).pin_memory()
d2 = torch.zeros(n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
out["duplex_GBps_each"] = 10 * n * 8 / (time.perf_counter() - t0) / 1e9

A = ab.CsrOperator.laplace2d(4096, 4096)
x = h.numpy()
y = h2.numpy()
x[:] = 1.0
for _ in range(2):
    A(x, y)
t0 = time.perf_counter()
for _ in range(5):
    A(x, y)
out["op_hostvec_ms"] = (time.perf_counter() - t0) / 5 * 1e3
r0 = ab.hashed_start_vector(A.n).cpu().numpy()
for hb in (True,):
    ab.solve(A, A.n, 10, 40, "LA", tol=0.0, mxiter=1, resid=r0, eupd=False, host_buffers=hb)
    t0 = time.perf_counter()
    r = ab.solve(A, A.n, 10, 40, "LA", tol=0.0, mxiter=1, resid=r0, eupd=False, host_buffers=hb)
    dt = time.perf_counter() - t0
    out["host_solve_ms_per_step"] = dt / int(r.iparam[8]) * 1e3
    out["host_solve_steps"] = int(r.iparam[8])
print(json.dumps(out))
