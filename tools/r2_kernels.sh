#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/r2_kernel_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_kernel_tests.log
for op in "laplace2d 4096" "laplace3d 256"; do set -- $op
  python tools/spmv_probe.py --op $1 --nx $2 --variants 0,4,1 2>/dev/null | tail -1
done
python - <<'PY'
# V*Q: tensor-core kernel vs the SIMT kernel, restart shapes of config 2 (40 -> ~14) and config 3 (64 -> ~30)
import json, os, subprocess, sys
sys.path.insert(0, ".")
for vq in ("mma", "simt"):
    env = dict(os.environ, AB200_VQ=vq)
    code = r'''
import sys, json
sys.path.insert(0, ".")
import arpack_ng_b200 as ab
L = ab.lib()
out = {}
for (n, ncv, kout) in ((1 << 24, 40, 14), (1 << 24, 64, 30), (1 << 24, 64, 21), (1 << 21, 64, 30)):
    L.ab200_kernel_probe_f64(n, ncv, ncv, 2, 2, kout)
    ab.profile(enable=True, reset=True)
    L.ab200_kernel_probe_f64(n, ncv, ncv, 12, 2, kout)
    p = ab.profile(enable=False)
    for k, v in p.items():
        if k.startswith("vq"):
            out[f"n={n} kin={ncv} kout={kout}"] = {"kernel": k, "ms": round(v["ms"] / v["launches"], 3), "GBps": round(v["bytes"] / v["ms"] / 1e6, 1)}
print(json.dumps(out))
'''
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(vq, r.stdout.strip()[-900:], r.stderr.strip()[-300:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_vq_mma" -c 2 -o gpurun_out/r2_ncu_vq_mma -f python -c "
import sys; sys.path.insert(0,'.')
import arpack_ng_b200 as ab
ab.lib().ab200_kernel_probe_f64(1<<24, 64, 64, 3, 2, 30)" > gpurun_out/r2_ncu_vq_mma.log 2>&1; echo "ncu rc=$?"
