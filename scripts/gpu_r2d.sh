#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/env_ab.py 0 BMPC_MAX_SOC=1 BMPC_MAX_SOC=0 BMPC_THREADS=160 BMPC_THREADS=192 > gpurun_out/env_ab.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/env_ab.log; tail -3 gpurun_out/bench.err; head -c 600 gpurun_out/bench_ref.json
