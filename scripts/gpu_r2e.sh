#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for cfg in exp1_1024 exp2_8192 exp1_N20_tight_8192 spec_mixed_65536; do
  timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err; echo "$cfg rc=$?"
done
BMPC_VEC_GLOBAL=1 timeout 900 python bench.py --config exp1_N20_tight_8192 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_exp1_N20_tight_8192_vecglobal.json 2> gpurun_out/bench_n20vg.err
timeout 1500 ncu --set full --clock-control none --profile-from-start off -k regex:k_solve -o gpurun_out/prof_n20 -f python scripts/profile_batch.py 8192 exp1_N20_tight_8192 > gpurun_out/prof_n20.log 2>&1
ncu -i gpurun_out/prof_n20.ncu-rep --page raw --csv > gpurun_out/prof_n20_raw.csv 2>/dev/null
rm -f gpurun_out/prof_n20.ncu-rep
python scripts/ncu_summary.py gpurun_out/prof_n20_raw.csv > gpurun_out/prof_n20_metrics.txt 2>&1
tail -2 gpurun_out/prof_n20.log
for f in gpurun_out/bench_*.json; do echo $f; head -c 300 $f; echo; done
