#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gram.py tests/test_gpu_parpack.py -x -q > gpurun_out/r2_gram_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_gram_tests.log
timeout 600 python tools/run_configs.py 5 > gpurun_out/r2_config5_n1.json 2> gpurun_out/r2_config5_n1.err; echo "config5 rc=$?"; tail -c 300 gpurun_out/r2_config5_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_config5_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('problem','info','nconv','restarts','nopx','time_to_solution_s','lanczos_steps_per_s')})
print(d['kernels'])
print('sigma', d.get('singular_values'))
PY
AB200_L2_WINDOW=0 timeout 600 python tools/run_configs.py 5 > gpurun_out/r2_config5_n1_nowindow.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_config5_n1_nowindow.json').read().strip().splitlines()[-1])
print('no L2 window:', d['time_to_solution_s'], {k:v for k,v in d['kernels'].items() if k.startswith('gram')})
PY
# ncu: gram SpMV kernels (config 5), then V*Q and the 7-point SpMV (config 3 at 256^3, ncv 64)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_rows -s 40 -c 4 -o gpurun_out/r2_ncu_gram -f python tools/run_configs.py 5 --mxiter 2 > gpurun_out/r2_ncu_gram.log 2>&1; echo "ncu gram rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_vq_tma|k_csr_spmv_bulk" -s 100 -c 6 -o gpurun_out/r2_ncu_vq_spmv -f python tools/run_configs.py 3 --scale 0.5 --mxiter 3 > gpurun_out/r2_ncu_vq.log 2>&1; echo "ncu vq rc=$?"
ls -la gpurun_out/*.ncu-rep
