#!/bin/bash
# round 2: everything that goes to profiles/ from ONE box with the final code (1 GPU)
mkdir -p gpurun_out/r2
O=gpurun_out/r2
V=boundmpc_b200/variants
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
for cfg in exp1_1024 exp2_8192 exp1_N20_tight_8192 spec_mixed_65536; do
  timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 > $O/bench_$cfg.json 2> $O/bench_$cfg.err; echo "$cfg rc=$?"
done
timeout 1200 python scripts/shard_sweep.py 0 1 2 3 4 5 6 7 > $O/shard_sweep.log 2>&1; cp gpurun_out/shard_sweep.json $O/shard_sweep.json; cut -c1-120 $O/shard_sweep.log
BMPC_LIB=$V/trace.so timeout 600 python scripts/trace_util.py 0 > $O/trace_util_s0.txt 2>&1
BMPC_LIB=$V/trace.so timeout 600 python scripts/trace_util.py 4 > $O/trace_util_s4.txt 2>&1; head -4 $O/trace_util_s4.txt
BMPC_LIB=$V/timing.so timeout 300 python scripts/phase_timing.py 148 > $O/phase_cycles_b148.txt 2>&1
BMPC_LIB=$V/timing.so timeout 300 python scripts/phase_timing.py 8192 > $O/phase_cycles_b8192.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
for w in 0 1; do
  if [ $w = 0 ]; then export BMPC_L2_WINDOW=1; else unset BMPC_L2_WINDOW; fi
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none --profile-from-start off -k regex:k_solve --csv \
     --log-file $O/dram_nowindow$w.csv python scripts/profile_batch.py 8192 > /dev/null 2>&1
done
unset BMPC_L2_WINDOW
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_solve -o $O/prof -f python scripts/profile_batch.py 8192 > $O/prof.log 2>&1
ncu -i $O/prof.ncu-rep --page source --csv > $O/prof_source.csv 2>/dev/null
ncu -i $O/prof.ncu-rep --page raw --csv > $O/prof_raw.csv 2>/dev/null
rm -f $O/prof.ncu-rep
timeout 1500 ncu --set full --clock-control none --profile-from-start off -k regex:k_solve -o $O/prof_n20 -f python scripts/profile_batch.py 8192 exp1_N20_tight_8192 > $O/prof_n20.log 2>&1
ncu -i $O/prof_n20.ncu-rep --page raw --csv > $O/prof_n20_raw.csv 2>/dev/null
rm -f $O/prof_n20.ncu-rep
timeout 600 python scripts/env_ab.py 0 BMPC_L2_WINDOW=1 BMPC_X=0 > $O/env_ab_l2.log 2>&1; cat $O/env_ab_l2.log
tail -n 3 $O/dram_nowindow0.csv $O/dram_nowindow1.csv | cut -c1-200
ls -la $O
