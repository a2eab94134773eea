#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python - <<'PY'
import json, os, subprocess, sys
code = r'''
import sys, json
sys.path.insert(0, ".")
import arpack_ng_b200 as ab
L = ab.lib()
out = {}
for (n, ncv, kout) in ((1 << 24, 40, 14), (1 << 24, 40, 11), (1 << 24, 64, 30), (1 << 24, 64, 21), (1 << 21, 64, 30), (1<<24, 30, 8), (1<<24, 48, 17)):
    L.ab200_kernel_probe_f64(n, ncv, ncv, 2, 2, kout)
    ab.profile(enable=True, reset=True)
    L.ab200_kernel_probe_f64(n, ncv, ncv, 12, 2, kout)
    p = ab.profile(enable=False)
    for k, v in p.items():
        if k.startswith("vq"):
            out[f"{n}/{ncv}/{kout}"] = (k, round(v["ms"] / v["launches"], 3), round(v["bytes"] / v["ms"] / 1e6))
print(json.dumps(out))
'''
r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
print("vq", r.stdout.strip()[-900:], r.stderr.strip()[-300:])
PY
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
tail -4 gpurun_out/r2_gputests.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo rc=$?
tail -c 600 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "rt/step", d.get("host_round_trips_per_lanczos_step"), "agg", d["roofline"]["lanczos_step_aggregate"])
print("kernels", {k:(v["ms"],v["GBps"]) for k,v in d["roofline"]["all_kernels"].items()})
for k in ("e2e","e2e_registered_host_csr","registered_op_mode","value_mxiter1","fullsize_parity","cpu_baseline","clocks"):
    print("  ",k, json.dumps(d.get(k))[:330])
c=d.get("config3") or {}
print("   config3", c.get("value"), c.get("step_hbm"), c.get("kernels"), c.get("error"))
PY
