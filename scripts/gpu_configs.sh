#!/bin/bash
# GPU tests + one bench line per BASELINE configuration (-> profiles/ via scripts/collect_profiles.py)
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -n "passed\|failed\|FAILED\|^E  " $O/pytest_gpu.log | head
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
for cfg in exp1_1024 exp2_8192 exp1_N20_tight_8192 spec_mixed_65536; do
  timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 > $O/bench_$cfg.json 2> $O/bench_$cfg.err; echo "$cfg rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('$O/bench*.json')):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1]); s=b.get('solver',{})
        print(f.split('/')[-1],'value %.0f e2e %.0f'%(b['value'],b['e2e']['value']),'succ',s.get('success'),'it',s.get('iters_mean'),s.get('iters_max_rank0'),s.get('status_hist_rank0'),'cpu',b.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
