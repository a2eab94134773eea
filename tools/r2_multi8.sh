#!/bin/bash
N=${1:-8}
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
export AB200_P2P_TIMEOUT_S=30
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
run tools/multigpu_check.py > gpurun_out/r2_multigpu_check_n${N}_fused.log 2>&1; echo "check fused rc=$?"
grep -E "MULTIGPU_CHECK|MISMATCH|rror" gpurun_out/r2_multigpu_check_n${N}_fused.log | head -12
for mode in fused unfused; do
  extra=""
  if [ $mode = unfused ]; then export AB200_FUSED_REDUCE=0 AB200_PEER_HALO=0; extra="--no-config3"; fi
  run bench.py --gpus $N --steps 5 --warmup 3 $extra > gpurun_out/r2_bench_n${N}_$mode.json 2> gpurun_out/r2_bench_n${N}_$mode.err; echo "bench $mode rc=$?"
  tail -c 300 gpurun_out/r2_bench_n${N}_$mode.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n${N}_$mode.json').read().strip().splitlines()[-1])
    print('$mode N=${N} value',round(d['value'],1),'ms/lanczos step',round(d['ms_per_lanczos_step'],4),'launches',d['gpu_launches'],'allreduces',d['allreduces'],'agg',d['roofline']['lanczos_step_aggregate'])
    print('  kernels', {k:(v['ms'],v['GBps']) for k,v in d['roofline']['all_kernels'].items()})
    print('  registered', (d.get('registered_op_mode') or {}).get('value'))
    c=d.get('config3') or {}
    print('  config3', c.get('value'), c.get('step_hbm'), c.get('error'))
except Exception as e:
    print('unreadable', e)
PY
done
