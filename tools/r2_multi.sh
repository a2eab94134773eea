#!/bin/bash
# usage: r2_multi.sh N [bench-args...]   -- multi-GPU parity against the oracle's logical-rank mode in three transport modes
# (fused peer-memory reductions + peer halo, stand-alone peer all-reduce + NCCL halo, NCCL only), then the bench line
N=$1; shift
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
export AB200_P2P_TIMEOUT_S=30
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
run tools/multigpu_check.py > gpurun_out/r2_multigpu_check_n${N}_fused.log 2>&1; echo "check fused rc=$?"
grep -E "MULTIGPU_CHECK|MISMATCH|rror" gpurun_out/r2_multigpu_check_n${N}_fused.log | head -12
AB200_FUSED_REDUCE=0 AB200_PEER_HALO=0 run tools/multigpu_check.py > gpurun_out/r2_multigpu_check_n${N}_p2p.log 2>&1; echo "check p2p rc=$?"
grep -E "MULTIGPU_CHECK|MISMATCH|rror" gpurun_out/r2_multigpu_check_n${N}_p2p.log | head -12
AB200_P2P=0 run tools/multigpu_check.py > gpurun_out/r2_multigpu_check_n${N}_nccl.log 2>&1; echo "check nccl rc=$?"
grep -E "MULTIGPU_CHECK|MISMATCH|rror" gpurun_out/r2_multigpu_check_n${N}_nccl.log | head -12
for mode in fused unfused; do
  if [ $mode = unfused ]; then export AB200_FUSED_REDUCE=0 AB200_PEER_HALO=0; fi
  run bench.py --gpus $N --steps 5 --warmup 3 --no-config3 "$@" > gpurun_out/r2_bench_n${N}_$mode.json 2> gpurun_out/r2_bench_n${N}_$mode.err; echo "bench $mode rc=$?"
  tail -c 400 gpurun_out/r2_bench_n${N}_$mode.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n${N}_$mode.json').read().strip().splitlines()[-1])
    print('$mode N=${N} value',round(d['value'],1),'ms/lanczos step',round(d['ms_per_lanczos_step'],4),'launches',d['gpu_launches'],'allreduces',d['allreduces'],'agg',d['roofline']['lanczos_step_aggregate'])
    print('  registered', (d.get('registered_op_mode') or {}).get('value'))
except Exception as e:
    print('unreadable', e)
PY
done
