#!/usr/bin/env python
"""Per-kernel bandwidth of the complex Arnoldi path (znaupd_c) on a large synthetic operator.

1-D complex convection-diffusion stencil on n points (default 4 Mi complex128 = 64 MiB per vector, V = n x ncv), applied on
the device with three fused torch element-wise ops (the OP belongs to the caller); znaupd_c with a fixed restart budget.
Prints one JSON line: Arnoldi steps/s and, per library kernel, launches / ms / algorithmic GB/s (CUDA events on the
launching stream, bytes as charged by the library: complex element = 16 B)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4 * 1024 * 1024)
    ap.add_argument("--nev", type=int, default=6)
    ap.add_argument("--ncv", type=int, default=30)
    ap.add_argument("--restarts", type=int, default=4)
    args = ap.parse_args()
    import torch
    import arpack_ng_b200 as ab
    n = args.n
    h = 1.0 / (n + 1)
    rho = 10.0
    dd = torch.full((n,), 2.0, dtype=torch.complex128, device="cuda") + 1j * torch.linspace(0, 1, n, dtype=torch.float64,
                                                                                             device="cuda")
    lo, up = -1.0 - rho * h / 2, -1.0 + rho * h / 2

    def op(x, y, *_):
        torch.mul(dd, x, out=y)
        y[1:].add_(x[:-1], alpha=lo)
        y[:-1].add_(x[1:], alpha=up)
    rng = np.random.default_rng(0)
    r0 = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    ab.solve_complex(op, n, args.nev, args.ncv, "LM", tol=1e-10, mxiter=1, resid=r0, eupd=False)   # warm-up
    torch.cuda.synchronize()
    ab.profile(enable=True, reset=True)
    t0 = time.perf_counter()
    r = ab.solve_complex(op, n, args.nev, args.ncv, "LM", tol=1e-10, mxiter=args.restarts, resid=r0, eupd=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    prof = ab.profile(enable=False)
    out = {"n": n, "nev": args.nev, "ncv": args.ncv, "dtype": "complex128", "restarts": int(r.iparam[2]),
           "nopx": int(r.iparam[8]), "info": int(r.info), "seconds": dt, "arnoldi_steps_per_s": int(r.iparam[8]) / dt,
           "kernels": {k: {"launches": v["launches"], "ms": round(v["ms"], 3),
                           "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None}
                       for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
