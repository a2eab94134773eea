#!/bin/bash
# round 2, first GPU pass: whole GPU suite, then the bench line with and without device-resident sweeps
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
tail -5 gpurun_out/r2_gputests.log
AB200_DEFER=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-registered --no-config3 --no-extras > gpurun_out/r2_bench_defer0.json 2> gpurun_out/r2_bench_defer0.err; echo rc=$?
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo rc=$?
tail -c 1500 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_defer0.json","gpurun_out/r2_bench_n1.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],1), "rt/step", d.get("host_round_trips_per_lanczos_step"), "agg", d["roofline"]["lanczos_step_aggregate"])
        for k in ("e2e","e2e_registered_host_csr","registered_op_mode","value_mxiter1","fullsize_parity","cpu_baseline"):
            print("  ",k, json.dumps(d.get(k))[:400])
        print("   config3", json.dumps(d.get("config3"))[:1500])
    except Exception as e:
        print(f, "unreadable", e)
PY
