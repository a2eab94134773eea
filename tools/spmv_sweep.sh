#!/bin/bash
# SpMV variants on config 2's operator: registered-mode bench per variant (the fused kernel) + RCI (plain kernel)
for v in 0 1 2 3 4 5; do
  AB200_SPMV_BULK=$v timeout 300 python bench.py --no-cpu --no-e2e --no-registered --steps 3 --op-mode registered 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['all_kernels']; print('bulk variant $v registered', round(d['value'],1), {n:k[n] for n in k if 'spmv' in n})"
done
AB200_SPMV=stream timeout 300 python bench.py --no-cpu --no-e2e --no-registered --steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['all_kernels']; print('stream rci', round(d['value'],1), {n:k[n] for n in k if 'spmv' in n})"
timeout 300 python bench.py --no-cpu --no-e2e --no-registered --steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['all_kernels']; print('bulk rci', round(d['value'],1), {n:k[n] for n in k if 'spmv' in n})"
