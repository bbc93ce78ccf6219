#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -n "passed\|failed\|FAILED\|^E  " gpurun_out/pytest_gpu.log | head -30
if [ -n "$AB" ]; then timeout 600 python scripts/env_ab.py 0 $AB > gpurun_out/env_ab.log 2>&1; cat gpurun_out/env_ab.log; fi
