#!/bin/bash
# One gpurun call: smoke, GPU parity tests, bench, ncu launch list, ncu full capture of k_solve.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 2 -c 1 -o gpurun_out/prof_k_solve \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
