"""Where does the time between kernels go?  (round 2: kernel time was 0.88 of the elapsed time of a solve although the
host no longer synchronises per step.)

  a. back-to-back orthogonalisation steps (ab200_kernel_probe_f64, no host sync inside): wall time of the loop against
     the sum of the per-launch CUDA-event times of the same loop run under the library's profiler;
  b. sustained copy bandwidth in 0.25 s windows for a few seconds (does the power cap pull it down over time?).
One JSON line."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arpack_ng_b200 as ab  # noqa: E402

L = ab.lib()
out = {}
n, ncv = 4096 * 4096, 40
for j in (30,):
    res = {}
    L.ab200_kernel_probe_f64(n, j, ncv, 2, 0, 1)     # warm-up (allocation paths, attributes)
    for iters in (0, 300):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        L.ab200_kernel_probe_f64(n, j, ncv, iters, 0, 1)
        torch.cuda.synchronize()
        res[f"wall_s_iters{iters}"] = time.perf_counter() - t0
    ab.profile(enable=True, reset=True)
    L.ab200_kernel_probe_f64(n, j, ncv, 300, 0, 1)
    torch.cuda.synchronize()
    prof = ab.profile(enable=False)
    res["profiled_ms_sum"] = sum(v["ms"] for v in prof.values())
    res["profiled"] = {k: round(v["ms"], 2) for k, v in prof.items()}
    res["unprofiled_loop_ms"] = 1e3 * (res["wall_s_iters300"] - res["wall_s_iters0"])
    res["ratio_unprofiled_over_profiled"] = res["unprofiled_loop_ms"] / res["profiled_ms_sum"]
    out[f"orth_j{j}"] = res

# b. sustained copy bandwidth
a = torch.empty(1 << 28, dtype=torch.float64, device="cuda")   # 2 GiB
b = torch.empty_like(a)
torch.cuda.synchronize()
win = []
t_end = time.perf_counter() + 4.0
while time.perf_counter() < t_end:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        b.copy_(a)
    e1.record()
    e1.synchronize()
    win.append(round(4 * 2 * a.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1))
out["copy_GBps_windows"] = win
print(json.dumps(out))
