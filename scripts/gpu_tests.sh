#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -n "passed\|failed\|FAILED\|^E  " gpurun_out/pytest_gpu.log | head -30
timeout 600 python scripts/env_ab.py 4 BMPC_HARD_CONTINUE=1 BMPC_HARD_CONTINUE=0 > gpurun_out/env_ab_hc.log 2>&1; cat gpurun_out/env_ab_hc.log
