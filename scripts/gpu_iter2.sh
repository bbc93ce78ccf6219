#!/bin/bash
# Development iteration on the GPU box: parity tests, throughput of the default launch shape, per-phase cycles.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python scripts/sweep_ctas.py 8192 > gpurun_out/sweep_ctas.log 2>&1
cat gpurun_out/sweep_ctas.log
if [ -f boundmpc_b200/libboundmpc_b200_timing.so ]; then
BMPC_LIB=boundmpc_b200/libboundmpc_b200_timing.so timeout 600 python scripts/phase_timing.py 148 > gpurun_out/phase_b148.txt 2>&1
cat gpurun_out/phase_b148.txt
fi
