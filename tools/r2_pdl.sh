#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_parpack.py -x -q 2>&1 | tail -2
for pdl in 1 0 1 0; do
AB200_PDL=$pdl timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-config3 --no-extras > gpurun_out/r2_bench_pdl$pdl.json 2> gpurun_out/r2_bench_pdl$pdl.err; echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_pdl$pdl.json").read().strip().splitlines()[-1])
print("PDL=$pdl value", round(d["value"],1), "reg", round(d["registered_op_mode"]["value"],1), "ms/step", round(d["ms_per_lanczos_step"],4), "share", round(d["roofline"]["lanczos_step_aggregate"]["kernel_time_share_of_elapsed"],4))
PY
done
