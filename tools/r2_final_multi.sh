#!/bin/bash
# usage: r2_final_multi.sh N   -- official multi-GPU capture: parity against the oracle's logical ranks (fused peer-memory
# path; NCCL path for N <= 4), bench line at N GPUs (config 2 + config-3 block + time to solution), config 5 row-sharded
N=$1
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
export AB200_P2P_TIMEOUT_S=60
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
run tools/multigpu_check.py > gpurun_out/r2_multigpu_check_n${N}_fused.log 2>&1; echo "check fused rc=$?"
grep -E "MULTIGPU_CHECK|MISMATCH|rror" gpurun_out/r2_multigpu_check_n${N}_fused.log | cut -c1-200 | head -8
if [ "$N" -le 4 ]; then
  AB200_P2P=0 run tools/multigpu_check.py > gpurun_out/r2_multigpu_check_n${N}_nccl.log 2>&1; echo "check nccl rc=$?"
  grep -E "MULTIGPU_CHECK|MISMATCH|rror" gpurun_out/r2_multigpu_check_n${N}_nccl.log | cut -c1-200 | head -8
fi
run bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench_n${N}.err
run tools/run_configs.py 5 > gpurun_out/r2_config5_n${N}.json 2> gpurun_out/r2_config5_n${N}.err; echo "config5 rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n${N}.json').read().strip().splitlines()[-1])
    print('N=${N} value',round(d['value'],1),'ms/lanczos step',round(d['ms_per_lanczos_step'],4),'launches',d['gpu_launches'],'allreduces',d['allreduces'],'agg',d['roofline']['lanczos_step_aggregate'])
    print('  kernels', {k:(v['ms'],v['GBps']) for k,v in d['roofline']['all_kernels'].items()})
    print('  registered', (d.get('registered_op_mode') or {}).get('value'))
    c=d.get('config3') or {}
    print('  config3', c.get('value'), c.get('step_hbm'), c.get('time_to_solution'), c.get('error'))
    x=json.loads(open('gpurun_out/r2_config5_n${N}.json').read().strip().splitlines()[-1])
    print('  config5', {k:x[k] for k in ('info','nconv','restarts','nopx','time_to_solution_s','lanczos_steps_per_s')}, {k:v for k,v in x['kernels'].items() if v['ms']>20})
except Exception as e:
    print('unreadable', e)
PY
