#!/bin/bash
# usage: tools/scale_run.sh N  -- bench.py on N GPUs for config 2 (default) and config 3 (laplace3d 512^3), JSON lines into gpurun_out/
N=$1
run() { # args: tag, extra bench args
  tag=$1; shift
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --no-cpu --no-e2e "$@" > gpurun_out/scale_${tag}_n$N.json 2> gpurun_out/scale_${tag}_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --no-cpu --no-e2e "$@" > gpurun_out/scale_${tag}_n$N.json 2> gpurun_out/scale_${tag}_n$N.err
  fi
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/scale_${tag}_n$N.json") if l.startswith("{")][-1])
    print("${tag}", $N, "steps/s", round(d["value"], 2), "ms/lanczos step", round(d["ms_per_lanczos_step"], 3), "allreduces", d["allreduces"],
          "kernel share", round(d["roofline"]["lanczos_step_aggregate"]["kernel_time_share_of_elapsed"], 3),
          "agg frac", round(d["roofline"]["lanczos_step_aggregate"]["frac_of_peak"], 3))
except Exception as e:
    print("${tag}", $N, "FAILED", e)
PY
  tail -2 gpurun_out/scale_${tag}_n$N.err | cut -c1-300
}
run c2 --steps 3 --warmup 3
run c3 --workload laplace3d --nx 512 --restarts 6 --steps 2 --warmup 3
