#!/bin/bash
mkdir -p gpurun_out
for shape in lat thr; do
  if [ $shape = thr ]; then export BMPC_NO_LATENCY_SHAPE=1; fi
  for tool in memcheck racecheck; do
    timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_batch.py > gpurun_out/san_${shape}_$tool.log 2>&1
    echo "== $shape $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^N=|Error|hazard" gpurun_out/san_${shape}_$tool.log | head -8
  done
done
