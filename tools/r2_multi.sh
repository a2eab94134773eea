#!/bin/bash
# usage: r2_multi.sh N [bench-args...]   -- multi-GPU parity (peer-memory and NCCL reductions) + bench line at N GPUs
N=$1; shift
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
export AB200_P2P_TIMEOUT_S=60
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
run tools/multigpu_check.py > gpurun_out/r2_multigpu_check_n${N}_p2p.log 2>&1; echo "check p2p rc=$?"
grep -E "MULTIGPU_CHECK|MISMATCH|Error|error" gpurun_out/r2_multigpu_check_n${N}_p2p.log | head -20
AB200_P2P=0 run tools/multigpu_check.py > gpurun_out/r2_multigpu_check_n${N}_nccl.log 2>&1; echo "check nccl rc=$?"
grep -E "MULTIGPU_CHECK|MISMATCH|Error|error" gpurun_out/r2_multigpu_check_n${N}_nccl.log | head -20
run bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_n${N}.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n${N}.json').read().strip().splitlines()[-1])
    print('N=${N} value',round(d['value'],1),'ms/lanczos step',round(d['ms_per_lanczos_step'],4),'rt/step',d['host_round_trips_per_lanczos_step'],'agg',d['roofline']['lanczos_step_aggregate'],'path',d['allreduce_path'])
    print('  kernels', {k:(v['ms'],v['GBps']) for k,v in d['roofline']['all_kernels'].items()})
    print('  registered', json.dumps(d.get('registered_op_mode'))[:300])
    c=d.get('config3') or {}
    print('  config3', c.get('value'), c.get('step_hbm'), c.get('error'))
except Exception as e:
    print('unreadable', e)
PY
