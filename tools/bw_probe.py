"""HBM copy bandwidth on this box: burst (best of 10) vs sustained (3 s loop), the same method MEASURED_PEAKS.json
documents (torch b.copy_(a), read+write bytes, CUDA events).  Explains the gap between kernels timed alone (ncu) and
kernels timed inside a long solve."""
import json
import subprocess
import time

import torch

n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
b = torch.empty_like(a)
nbytes = 2 * n * 2
for _ in range(3):
    b.copy_(a)
torch.cuda.synchronize()
best = 0.0
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    time.sleep(0.05)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time(); k = 0
e0.record()
while time.time() - t0 < 3.0:
    for _ in range(20):
        b.copy_(a)
    k += 20
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus = k * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
# read-only stream: sum of a large tensor
x = torch.empty(1 << 29, dtype=torch.float64, device="cuda").zero_()
for _ in range(3):
    x.sum()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    x.sum()
e1.record(); torch.cuda.synchronize()
rd = 20 * x.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9
q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap",
                    "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
print(json.dumps({"copy_burst_GBps": best, "copy_sustained_GBps": sus, "read_only_sum_GBps": rd, "smi_after": q}))
