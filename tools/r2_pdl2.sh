#!/bin/bash
N=${1:-2}
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
export AB200_P2P_TIMEOUT_S=30
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
run tools/multigpu_check.py > gpurun_out/r2_multigpu_check_n${N}_pdl.log 2>&1; echo "check rc=$?"
grep -E "MULTIGPU_CHECK|MISMATCH|rror" gpurun_out/r2_multigpu_check_n${N}_pdl.log | cut -c1-200 | head -5
for pdl in 1 0 1 0; do
AB200_PDL=$pdl run bench.py --gpus $N --steps 5 --warmup 3 --no-config3 --no-extras > gpurun_out/r2_bench_n${N}_pdl$pdl.json 2> /dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n${N}_pdl$pdl.json").read().strip().splitlines()[-1])
print("N=$N PDL=$pdl value", round(d["value"],1), "reg", round(d["registered_op_mode"]["value"],1), "ms/step", round(d["ms_per_lanczos_step"],4))
PY
done
