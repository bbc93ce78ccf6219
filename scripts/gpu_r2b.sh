#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/slice_sweep.py 0 > gpurun_out/slice_sweep.log 2>&1
timeout 600 python scripts/shard_sweep.py 0 2 3 > gpurun_out/shard_sweep.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/slice_sweep.log; cut -c1-260 gpurun_out/shard_sweep.log
