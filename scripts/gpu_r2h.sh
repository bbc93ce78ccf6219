#!/bin/bash
mkdir -p gpurun_out
V=boundmpc_b200/variants
timeout 1200 python scripts/shard_sweep.py 0 1 2 3 4 5 6 7 > gpurun_out/shard_sweep.log 2>&1
cut -c1-230 gpurun_out/shard_sweep.log
BMPC_LIB=$V/trace.so timeout 600 python scripts/trace_util.py 4 > gpurun_out/trace_util_s4.log 2>&1; cat gpurun_out/trace_util_s4.log
timeout 600 python scripts/env_ab.py 4 BMPC_X=0 BMPC_NO_L2_WINDOW=1 > gpurun_out/env_ab_l2.log 2>&1; cat gpurun_out/env_ab_l2.log
