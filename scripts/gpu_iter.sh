#!/bin/bash
# Development iteration on the GPU box: parity tests, launch-shape sweep, optional ncu capture of the batch launch.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python scripts/sweep_variants.py 8192 > gpurun_out/sweep.log 2>&1
cat gpurun_out/sweep.log
if [ -n "$PROFILE" ]; then
  BMPC_THREADS=${PT:-256} BMPC_CTAS_PER_SM=${PC:-1} timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -o gpurun_out/prof_batch -f python scripts/profile_batch.py 8192 > gpurun_out/prof_batch.log 2>&1
  tail -2 gpurun_out/prof_batch.log
fi
