#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python scripts/sweep_ctas.py 8192 > gpurun_out/sweep_ctas.log 2>&1
cat gpurun_out/sweep_ctas.log
ncu --query-metrics 2>/dev/null | grep -i -E "icc|icache|inst_cache|gcc|l1i|ifetch|instruction" > gpurun_out/icc_metrics.txt
wc -l gpurun_out/icc_metrics.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_probe.json 2> gpurun_out/bench_probe.err
cat gpurun_out/bench_probe.json | cut -c1-400
