#!/bin/bash
mkdir -p gpurun_out
V=boundmpc_b200/variants
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/slice_sweep.py 0 4 5 6 > gpurun_out/slice_sweep.log 2>&1
BMPC_LIB=$V/trace.so timeout 600 python scripts/trace_util.py 0 > gpurun_out/trace_util.log 2>&1
BMPC_LIB=$V/timing.so timeout 300 python scripts/phase_timing.py 148 > gpurun_out/phase_b148.txt 2>&1
BMPC_LIB=$V/timing.so timeout 300 python scripts/phase_timing.py 8192 > gpurun_out/phase_b8192.txt 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_solve \
   -o gpurun_out/prof_r2c -f python scripts/profile_batch.py 8192 > gpurun_out/prof_r2c.log 2>&1
ncu -i gpurun_out/prof_r2c.ncu-rep --page source --csv > gpurun_out/prof_r2c_source.csv 2>/dev/null
ncu -i gpurun_out/prof_r2c.ncu-rep --page raw --csv > gpurun_out/prof_r2c_raw.csv 2>/dev/null
rm -f gpurun_out/prof_r2c.ncu-rep
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/slice_sweep.log gpurun_out/trace_util.log; head -3 gpurun_out/phase_b8192.txt; tail -2 gpurun_out/prof_r2c.log; ls -la gpurun_out | tail -8
