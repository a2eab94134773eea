#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for d in 1 0; do
AB200_DEFER=$d timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-registered --no-config3 --no-extras --profile-in-timed > gpurun_out/r2_bench_pit$d.json 2> gpurun_out/r2_bench_pit$d.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_pit$d.json').read().strip().splitlines()[-1])
print('defer=$d value',round(d['value'],1),'ms/solve',round(d['ms_per_step'],1),'pit',d['profile_in_timed'],'agg',d['roofline']['lanczos_step_aggregate'], d['clocks'])"
done
python tools/solve_timing.py 2>&1 | tail -2
