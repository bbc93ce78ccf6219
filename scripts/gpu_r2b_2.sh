#!/bin/bash
mkdir -p gpurun_out
SLICES=3,4,5,6 timeout 900 python scripts/shard_slice.py > gpurun_out/shard_slice.log 2>&1; cat gpurun_out/shard_slice.log | tail -12
