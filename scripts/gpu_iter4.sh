#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
echo "=== two-pass"; timeout 600 python scripts/sweep_ctas.py 8192 2>&1 | tail -4
echo "=== single pass"; BMPC_SINGLE_PASS=1 timeout 600 python scripts/sweep_ctas.py 8192 2>&1 | tail -4
