"""A/B of development switches read by bmpc_create (development aid): python scripts/env_ab.py <shard> [VAR=val[,VAR=val] ...]
Every spec gets its own solver handle; kernel time of the shard (best of 4), interleaved twice to see the noise."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches
B = 8192
shard = int(sys.argv[1])
specs = sys.argv[2:] or ["BMPC_MAX_SOC=1", "BMPC_MAX_SOC=0"]
s0 = default_solver()
TOL = {'b200': {'tol': float(os.environ.get('AB_TOL', '1e-5'))}}
x0, p = batches.make_batch(s0, ("exp1", "exp2"), shard * B, B, bound_scale=True)
xd, pd = torch.from_numpy(x0).cuda(), torch.from_numpy(p).cuda()
solvers = []
for sp in specs:
    kv = dict(a.split("=") for a in sp.split(",") if a)
    old = {k: os.environ.get(k) for k in kv}
    os.environ.update(kv)
    solvers.append(default_solver(solver_opts=TOL))
    for k, v in old.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
ref = None
for rnd in range(2):
    for sp, s in zip(specs, solvers):
        out = s.solve_batch(xd, pd); torch.cuda.synchronize()
        best = 1e9
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.solve_batch(xd, pd, out); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        it = out["iters"]
        xh = out["x"].cpu().numpy()
        if ref is None: ref = xh
        same = bool(np.array_equal(xh, ref))
        print(f"round {rnd} {sp:40s} {best:.2f} ms  {B / best * 1e3:.0f} solves/s  iters mean {it.double().mean():.3f} max {int(it.max())} ok {int((out['status'] == 0).sum())} bitwise same as first: {same}", flush=True)
