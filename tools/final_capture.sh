#!/bin/bash
# The round's official single-GPU capture: bench line, reference arm, ncu launch list and one ncu --set full pass.
# usage (on the GPU box): tools/final_capture.sh <tag>     -> files under gpurun_out/<tag>_*
tag=${1:-r1}
set -x
timeout 900 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err || exit 1
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
CMD="python bench.py --steps 1 --warmup 3 --restarts 1 --no-e2e --no-cpu --no-registered"
timeout 600 $CMD > /dev/null 2>&1 || exit 2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/${tag}_ncu_launches.csv $CMD > gpurun_out/${tag}_ncu_launches.log 2>&1
# one Lanczos step around j ~ 31 of the last (timed) solve: skip the 3 warm-up solves (3 x ~352 matching launches) + 29 steps
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_orth|k_upd|k_csr_spmv|k_vq|k_start" -s 1200 -c 10 -f -o gpurun_out/${tag}_prof_full $CMD > gpurun_out/${tag}_ncu_full.log 2>&1
tail -3 gpurun_out/${tag}_ncu_full.log
