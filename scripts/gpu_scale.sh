#!/bin/bash
# weak-scaling bench on N GPUs (the driver's torchrun launch line) next to the 1-GPU line and the reference arm of the same box
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
     > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n=$N rc=$?"
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_box$N.json 2> gpurun_out/bench_n1_box$N.err; echo "n=1 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_box$N.json 2> gpurun_out/bench_reference_box$N.err; echo "ref rc=$?"
for n in n$N n1_box$N reference_box$N; do python - <<PY
import json
try:
    b=json.loads(open('gpurun_out/bench_$n.json').read().strip().splitlines()[-1])
    print('$n', 'value %.0f e2e %.0f ms %.2f gather_ms %.3f' % (b['value'], b['e2e']['value'], b['ms_per_step'], b.get('gather_ms',0)), [round(r['kernel_ms'],2) for r in b.get('solver',{}).get('per_rank',[])], b.get('cpu_baseline',{}).get('cores'))
except Exception as e: print('$n', 'ERR', e)
PY
done
tail -3 gpurun_out/bench_n$N.err
