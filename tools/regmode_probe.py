import sys, time, os
sys.path.insert(0, ".")
import torch
import arpack_ng_b200 as ab
A = ab.CsrOperator.laplace2d(4096, 4096)
r0 = ab.hashed_start_vector(A.n)
def run(reg, tag):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = ab.solve(A, A.n, 10, 40, "LA", tol=1e-10, mxiter=4, resid=r0, eupd=False, registered_op=A if reg else None)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(tag, "reg" if reg else "rci", int(r.iparam[8]), "steps", round(dt, 3), "s", round(int(r.iparam[8]) / dt, 1), "steps/s", flush=True)
for i in range(3): run(False, "noprof")
for i in range(4): run(True, "noprof")
ab.profile(enable=True, reset=True)
for i in range(3): run(True, "prof")
for i in range(2): run(False, "prof")
p = ab.profile(enable=False)
print({k: (v["launches"], round(v["ms"], 1)) for k, v in p.items()})
for i in range(3): run(True, "noprof2")
for i in range(2): run(False, "noprof2")
