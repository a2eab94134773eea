#!/bin/bash
# What round 1 could not run once its GPU budget was spent -- one gpurun call at the start of the next round:
#   gpurun --timeout 1500 -- 'bash tools/next_round_capture.sh r2'
# 1. the whole GPU suite (every file has been run on a B200, but not in one session with the final binary);
# 2. the default bench line (value, registered-host-CSR e2e, hand-off e2e, cpu_baseline) and the reference arm;
# 3. ncu --set full of the kernels that have no capture yet: k_vq_tma (restart of the real path; the round-1 capture
#    window held no restart) and the complex kernels k_zdots / k_zupdate / k_zvq;
# 4. the complex probe.
# Files land under gpurun_out/<tag>_*; copy the summaries into profiles/.
tag=${1:-r2}
mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
timeout 120 python tools/complex_probe.py > gpurun_out/${tag}_complex_probe.json 2> gpurun_out/${tag}_complex_probe.err
# restart kernel of the real path: two restarts' worth of launches of the timed solve
CMD="python bench.py --steps 1 --warmup 3 --restarts 3 --no-e2e --no-cpu --no-registered"
timeout 300 $CMD > /dev/null 2>&1 || exit 2
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_vq" -s 9 -c 3 -f -o gpurun_out/${tag}_prof_vq $CMD > gpurun_out/${tag}_ncu_vq.log 2>&1
# complex kernels: a short probe (the kernels are launched ~250 times; take a few from the middle)
ZCMD="python tools/complex_probe.py --restarts 2"
timeout 120 $ZCMD > /dev/null 2>&1 || exit 3
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_zdots|k_zupdate|k_zvq" -s 60 -c 6 -f -o gpurun_out/${tag}_prof_cplx $ZCMD > gpurun_out/${tag}_ncu_cplx.log 2>&1
python tools/ncu_summarize.py gpurun_out/${tag}_prof_vq.ncu-rep gpurun_out/${tag}_ncu_vq_summary.json
python tools/ncu_summarize.py gpurun_out/${tag}_prof_cplx.ncu-rep gpurun_out/${tag}_ncu_cplx_summary.json
ls -la gpurun_out | tail -20
