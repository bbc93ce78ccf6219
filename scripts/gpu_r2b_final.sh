#!/bin/bash
# round 2, second session: everything that goes to profiles/r2b_* from ONE box with the final code (1 GPU).
# Defaults of this build: bench at the reference's tolerance (ipopt.tol = 1e-5) with the tight-tolerance block, pass A = 4 iterations.
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
V=boundmpc_b200/variants
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
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
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none --profile-from-start off -k regex:k_solve --csv \
     --log-file $O/dram.csv python scripts/profile_batch.py 8192 > /dev/null 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_solve -o $O/prof -f python scripts/profile_batch.py 8192 > $O/prof.log 2>&1
ncu -i $O/prof.ncu-rep --page source --csv > $O/prof_source.csv 2>/dev/null
ncu -i $O/prof.ncu-rep --page raw --csv > $O/prof_raw.csv 2>/dev/null
rm -f $O/prof.ncu-rep
tail -n 3 $O/dram.csv | cut -c1-200
python - <<PY
import json,glob
for f in sorted(glob.glob('$O/bench*.json')):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1]); s=b.get('solver',{})
        print(f.split('/')[-1],'value %.0f e2e %.0f'%(b['value'],b['e2e']['value']),'succ',s.get('success'),'it',s.get('iters_mean'),s.get('iters_max_rank0'),s.get('status_hist_rank0'),'cpu',b.get('cpu_baseline',{}).get('value'), 'tight', (b.get('tight_tol') or {}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
ls -la $O
