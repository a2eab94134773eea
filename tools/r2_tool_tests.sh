#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_arpackmm.py tests/test_mmio.py -m gpu -q --durations=8 > gpurun_out/r2_tool_tests.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|Error:|assert " gpurun_out/r2_tool_tests.log | cut -c1-300 | head -40
grep -A10 "slowest" gpurun_out/r2_tool_tests.log | head -12
