#!/bin/bash
# One gpurun call: A/B of library variants (throughput, outputs vs the first), per-phase cycles, GPU tests.
mkdir -p gpurun_out
V=boundmpc_b200/variants
if [ -n "$LIBS" ]; then timeout 600 python scripts/ab_variants.py 8192 $LIBS > gpurun_out/ab.log 2>&1; fi
for t in $TLIBS; do
  BMPC_LIB=$V/$t timeout 300 python scripts/phase_timing.py 148 > gpurun_out/phase_${t%.so}.txt 2>&1
done
if [ -n "$RUNTESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; fi
rm -f gpurun_out/ab_*.npz
cat gpurun_out/ab.log
tail -3 gpurun_out/pytest_gpu.log
