#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "vq" 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_gram.py -x -q 2>&1 | tail -1
timeout 600 python tools/run_configs.py 5 2>/dev/null | python -c "
import json,sys
x=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:x[k] for k in ('info','nconv','restarts','nopx','time_to_solution_s')}, {k:v for k,v in x['kernels'].items() if k.startswith('gram')})"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_vq_mma" -c 1 -o gpurun_out/r2_prof_vq -f python -c "
import sys; sys.path.insert(0,'.')
import arpack_ng_b200 as ab
ab.lib().ab200_kernel_probe_f64(1<<24, 64, 64, 2, 2, 30)" > gpurun_out/r2_ncu_vq.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json, os, subprocess, sys
code = r'''
import sys, json
sys.path.insert(0, ".")
import arpack_ng_b200 as ab
L = ab.lib()
out = {}
for (n, ncv, kout) in ((1 << 24, 40, 14), (1 << 24, 40, 11), (1 << 24, 64, 30), (1 << 24, 64, 21), (1 << 21, 64, 30), (1<<24, 48, 17)):
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
