#!/bin/bash
# ncu source-level capture of one k_solve launch (B = $1 instances, $2 CTAs per SM)
mkdir -p gpurun_out
BMPC_THREADS=128 BMPC_CTAS_PER_SM=${2:-1} timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -o gpurun_out/prof_src -f python scripts/profile_batch.py ${1:-1184} > gpurun_out/prof_src.log 2>&1
tail -2 gpurun_out/prof_src.log
ncu -i gpurun_out/prof_src.ncu-rep --page source --csv > gpurun_out/prof_src_source.csv 2>/dev/null
ncu -i gpurun_out/prof_src.ncu-rep --page raw --csv > gpurun_out/prof_src_raw.csv 2>/dev/null
ls -la gpurun_out/prof_src*
