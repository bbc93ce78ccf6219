#!/bin/bash
# GPU tests, then the default bench with and without the zero-copy paths of the host entry.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
for z in 1 0; do
  if [ $z = 1 ]; then export BMPC_NO_ZERO_COPY=1; else unset BMPC_NO_ZERO_COPY; fi
  timeout 300 python bench.py $BENCH_FLAGS > gpurun_out/bench_zc$z.json 2> gpurun_out/bench_zc$z.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_zc$z.json"))
print("no_zero_copy=$z", d["value"], d["e2e"], d["latency_b1_ms"]["p50"], d["latency_mpc_step_ms"]["p50"])
PY
done
