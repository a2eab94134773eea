"""Per-kernel bandwidth as a function of the column count j (kernel tuning aid, run on the GPU box).
usage: python tools/kernel_sweep.py [n] [ncv]"""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arpack_ng_b200 as ab

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16777216
ncv = int(sys.argv[2]) if len(sys.argv) > 2 else 40
L = ab.lib()
rows = []
for j in (1, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 48, 56, 64):
    if j > ncv:
        break
    L.ab200_kernel_probe_f64(n, j, ncv, 3, 0, 0)  # warm-up
    ab.profile(enable=True, reset=True)
    assert L.ab200_kernel_probe_f64(n, j, ncv, 10, 0, 0) == 0
    pr = ab.profile(enable=False)
    row = {"j": j}
    for k, v in pr.items():
        if v["launches"] >= 10 and v["ms"] > 0:
            row[k] = {"us": round(1e3 * v["ms"] / v["launches"], 1), "GBps": round(v["bytes"] / v["ms"] / 1e6, 0)}
    rows.append(row)
    print(json.dumps(row))
for kout in (11, 16, 21, ncv):
    L.ab200_kernel_probe_f64(n, ncv, ncv, 1, 2, kout)
    ab.profile(enable=True, reset=True)
    L.ab200_kernel_probe_f64(n, ncv, ncv, 5, 2, kout)
    pr = ab.profile(enable=False)
    print(json.dumps({"vq kout": kout, **{k: {"us": round(1e3 * v["ms"] / v["launches"], 1),
                                               "GBps": round(v["bytes"] / v["ms"] / 1e6, 0)} for k, v in pr.items()
                                          if k.startswith("vq")}}))
