"""Time the CSR SpMV variants alone on a BASELINE operator (default config 2: 2-D Laplacian 4096^2)."""
import argparse
import json
import sys

import torch

sys.path.insert(0, ".")
import arpack_ng_b200 as ab

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=4096)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--variants", default="0,1,2")
ap.add_argument("--op", default="laplace2d", choices=["laplace2d", "laplace3d", "convdiff"])
a = ap.parse_args()
L = ab.lib()
if a.op == "laplace2d":
    A = ab.CsrOperator.laplace2d(a.nx, a.nx)
elif a.op == "laplace3d":
    A = ab.CsrOperator.laplace3d(a.nx, a.nx, a.nx)
else:
    A = ab.CsrOperator.convdiff2d(a.nx, 100.0)
x = ab.hashed_start_vector(A.n)
y = torch.empty_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for v in [int(t) for t in a.variants.split(",")]:
    L.ab200_set_spmv_variant(v)
    for _ in range(3):
        A(x, y)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(a.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        A(x, y)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / a.reps
    out[v] = {"ms": ms, "GBps": A.spmv_bytes() / ms / 1e6}
L.ab200_set_spmv_variant(0)
print(json.dumps(out))
