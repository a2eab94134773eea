#!/bin/bash
# Round 2, official single-GPU capture: GPU suite, smoke, bench line, reference arm, ncu launch list, ncu --set full of
# the step kernels and of V*Q, time-to-solution runs of configs 4 and 5.   usage (GPU box): tools/r2_final_n1.sh
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
tail -3 gpurun_out/r2_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
timeout 1500 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
CMD="python bench.py --steps 1 --warmup 3 --restarts 1 --no-e2e --no-cpu --no-registered --no-config3 --no-extras"
timeout 600 $CMD > /dev/null 2>&1 || echo "CMD failed"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/r2_ncu_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
# one Lanczos step around j ~ 31 of the timed solve: 3 warm-up solves x ~352 matching launches + 29 steps x 5
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_orth|k_upd|k_csr_spmv|k_start" -s 1200 -c 10 -f -o gpurun_out/r2_prof_full $CMD > gpurun_out/r2_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_vq" -s 6 -c 1 -f -o gpurun_out/r2_prof_vq $CMD > gpurun_out/r2_ncu_vq.log 2>&1
for c in 4 5; do timeout 600 python tools/run_configs.py $c >> gpurun_out/r2_configs_time_to_solution.jsonl 2>> gpurun_out/r2_configs.err; done
tail -c 2000 gpurun_out/r2_configs_time_to_solution.jsonl | cut -c1-900
ls -la gpurun_out/r2_prof*.ncu-rep
