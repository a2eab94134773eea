#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
tail -3 gpurun_out/r2_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_val.json 2> gpurun_out/r2_bench_val.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_val.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "reg", round(d["registered_op_mode"]["value"],1), "agg", d["roofline"]["lanczos_step_aggregate"], d["clocks"])
c=d.get("config3") or {}
print("config3", c.get("value"), c.get("step_hbm"), c.get("error"))
PY
