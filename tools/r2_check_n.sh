#!/bin/bash
# usage: r2_check_n.sh N   -- tools/multigpu_check.py with the fused peer-memory path and with NCCL only
N=${1:-2}
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
export AB200_P2P_TIMEOUT_S=30
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
run tools/multigpu_check.py > gpurun_out/r2_multigpu_check_n${N}_fused.log 2>&1; echo "check fused rc=$?"
grep -E "MULTIGPU_CHECK|MISMATCH|rror|FP32" gpurun_out/r2_multigpu_check_n${N}_fused.log | cut -c1-260 | head -8
AB200_P2P=0 run tools/multigpu_check.py > gpurun_out/r2_multigpu_check_n${N}_nccl.log 2>&1; echo "check nccl rc=$?"
grep -E "MULTIGPU_CHECK|MISMATCH|rror" gpurun_out/r2_multigpu_check_n${N}_nccl.log | cut -c1-260 | head -8
