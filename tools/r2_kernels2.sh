#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -x -q > gpurun_out/r2_kernel_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_kernel_tests.log
python tools/spmv_probe.py --op laplace3d --nx 256 --variants 0,4 2>/dev/null | tail -1
AB200_SPMV_OWN=2 AB200_SPMV_BULK=2 python tools/spmv_probe.py --op laplace3d --nx 256 --variants 0 2>/dev/null | tail -1
AB200_SPMV_OWN=2 python tools/spmv_probe.py --op laplace3d --nx 256 --variants 0 2>/dev/null | tail -1
python - <<'PY'
import json, os, subprocess, sys
for vq in ("mma", "simt"):
    env = dict(os.environ, AB200_VQ=vq)
    code = r'''
import sys, json
sys.path.insert(0, ".")
import arpack_ng_b200 as ab
L = ab.lib()
out = {}
for (n, ncv, kout) in ((1 << 24, 40, 14), (1 << 24, 40, 11), (1 << 24, 64, 30), (1 << 24, 64, 21), (1 << 21, 64, 30), (1<<24, 30, 8)):
    L.ab200_kernel_probe_f64(n, ncv, ncv, 2, 2, kout)
    ab.profile(enable=True, reset=True)
    L.ab200_kernel_probe_f64(n, ncv, ncv, 12, 2, kout)
    p = ab.profile(enable=False)
    for k, v in p.items():
        if k.startswith("vq"):
            out[f"{n}/{ncv}/{kout}"] = (k, round(v["ms"] / v["launches"], 3), round(v["bytes"] / v["ms"] / 1e6))
print(json.dumps(out))
'''
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(vq, r.stdout.strip()[-700:], r.stderr.strip()[-300:])
PY
