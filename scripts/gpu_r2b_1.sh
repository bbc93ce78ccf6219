#!/bin/bash
# session 2 of round 2, call 1: GPU tests, bench at the reference tolerance (+ tight block), pass-A length at tol 1e-5
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tol5.json 2> gpurun_out/bench_tol5.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_tol5.json"))
    print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d["solver"]["iters_mean"], d["solver"]["success"], d["tight_tol"], d.get("latency_step_dropin_ms"), d["latency_b1_ms"], d.get("cpu_baseline"))
except Exception as e: print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_tol5.err
AB_TOL=1e-5 timeout 300 python scripts/slice_sweep.py 0 4 5 6 7 8 10 > gpurun_out/slice_tol5.log 2>&1; cat gpurun_out/slice_tol5.log
BMPC_SINGLE_PASS=1 AB_TOL=1e-5 timeout 120 python scripts/slice_sweep.py 0 6 >> gpurun_out/slice_tol5.log 2>&1; tail -1 gpurun_out/slice_tol5.log
