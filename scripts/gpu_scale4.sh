#!/bin/bash
mkdir -p gpurun_out
for n in 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 \
     > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "n=$n rc=$?"
done
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1c.json 2> gpurun_out/bench_n1c.err
for n in 1c 2 4; do python - <<PY
import json
try:
    b=json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
    print('$n', 'value %.0f e2e %.0f ms %.2f gather_ms %.3f' % (b['value'], b['e2e']['value'], b['ms_per_step'], b.get('gather_ms',0)), [round(r['kernel_ms'],2) for r in b['solver']['per_rank']])
except Exception as e: print('$n', 'ERR', e)
PY
done
