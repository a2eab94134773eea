#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
PROBE_ONLY=${PROBE_ONLY:-} timeout 600 python tools/arpackmm_probe.py > gpurun_out/r2_tool_probe.log 2>&1; echo "rc=$?"
cut -c1-330 gpurun_out/r2_tool_probe.log | tail -40
