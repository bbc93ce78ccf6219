#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 \
     > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "n=8 rc=$?"
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; echo "n=1 rc=$?"
for n in 1b 8; do python - <<PY
import json
try:
    b=json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
    print('$n', 'value %.0f e2e %.0f ms %.2f gather_ms %.3f' % (b['value'], b['e2e']['value'], b['ms_per_step'], b.get('gather_ms',0)), [round(r['kernel_ms'],2) for r in b['solver']['per_rank']])
except Exception as e: print('$n', 'ERR', e)
PY
done
tail -3 gpurun_out/bench_n8.err
