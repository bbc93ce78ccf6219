#!/bin/bash
# Development iteration: parity tests on the default library, then throughput + phase cycles of every variant library given.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for V in "$@"; do
  echo "=== variant $V"
  BMPC_LIB=boundmpc_b200/libboundmpc_b200$V.so timeout 600 python scripts/sweep_ctas.py 8192 2>&1 | tail -4
  if [ -f boundmpc_b200/libboundmpc_b200${V}_timing.so ]; then
    BMPC_LIB=boundmpc_b200/libboundmpc_b200${V}_timing.so timeout 600 python scripts/phase_timing.py 148 > gpurun_out/phase_b148$V.txt 2>&1
    cat gpurun_out/phase_b148$V.txt
  fi
done
