#!/bin/bash
# round 2, call A: GPU tests, bench, shard sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 1200 python scripts/shard_sweep.py 0 1 2 3 4 5 6 7 > gpurun_out/shard_sweep.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/shard_sweep.log | cut -c1-300; head -c 1500 gpurun_out/bench.json
