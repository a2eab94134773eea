"""Per-kernel fixed costs: the three sweeps of a Lanczos step on ONE GPU at the per-GPU size of config 2 on 8 GPUs
(n = 2^21 rows) against larger n, per-launch CUDA-event times and the time of the same loop without the profiler."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import arpack_ng_b200 as ab

L = ab.lib()
out = {}
for n in (1 << 21, 1 << 22, 1 << 24):
    for j in (25,):
        L.ab200_kernel_probe_f64(n, j, 40, 3, 0, 1)
        ab.profile(enable=True, reset=True)
        L.ab200_kernel_probe_f64(n, j, 40, 200, 0, 1)
        p = ab.profile(enable=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); L.ab200_kernel_probe_f64(n, j, 40, 0, 0, 1); torch.cuda.synchronize(); t_base = time.perf_counter() - t0
        t0 = time.perf_counter(); L.ab200_kernel_probe_f64(n, j, 40, 400, 0, 1); torch.cuda.synchronize(); t_loop = time.perf_counter() - t0
        rec = {k: {"us": round(1e3 * v["ms"] / v["launches"], 2), "GBps": round(v["bytes"] / v["ms"] / 1e6),
                   "ideal_us_at_6546": round(v["bytes"] / v["launches"] / 6546e3, 2)} for k, v in p.items() if not k.startswith("(")}
        rec["unprofiled_us_per_orth_step"] = round(1e6 * (t_loop - t_base) / 400, 2)
        out[f"n={n} j={j}"] = rec
print(json.dumps(out))
