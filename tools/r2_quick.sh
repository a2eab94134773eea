#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python tools/solve_timing.py 2>&1 | tail -2 | cut -c1-700
python - <<'PY'
import json, os, subprocess, sys
code = r'''
import sys, json
sys.path.insert(0, ".")
import arpack_ng_b200 as ab
L = ab.lib()
out = {}
for (n, ncv, kout) in ((1 << 24, 40, 14), (1 << 24, 40, 11), (1 << 24, 64, 30), (1 << 24, 64, 21), (1 << 21, 64, 30), (1<<24, 30, 8), (1<<24, 48, 17), (1<<24, 64, 64)):
    L.ab200_kernel_probe_f64(n, ncv, ncv, 2, 2, kout)
    ab.profile(enable=True, reset=True)
    L.ab200_kernel_probe_f64(n, ncv, ncv, 12, 2, kout)
    p = ab.profile(enable=False)
    for k, v in p.items():
        if k.startswith("vq") and v["ms"] > 0.5:
            out[f"{n}/{ncv}/{kout}"] = (k, round(v["ms"] / v["launches"], 3), round(v["bytes"] / v["ms"] / 1e6))
print(json.dumps(out))
'''
r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
print("vq", r.stdout.strip()[-900:], r.stderr.strip()[-300:])
PY
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "vq" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-config3 --no-extras > gpurun_out/r2_bench_quick.json 2> gpurun_out/r2_bench_quick.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_quick.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "registered", round(d["registered_op_mode"]["value"],1), "agg", d["roofline"]["lanczos_step_aggregate"], d["clocks"])
print("kernels", {k:(v["ms"],v["GBps"]) for k,v in d["roofline"]["all_kernels"].items()})
PY
