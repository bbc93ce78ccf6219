"""Kernel time of every shard of the 65,536-instance mixed workload on ONE GPU next to its ideal packed time
(development aid): python scripts/shard_sweep.py [shards...].  Ideal = total iterations x time per iteration-slot of a
tail-free launch (measured on the same shard sorted longest-first) / resident CTAs."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches
B = 8192
shards = [int(a) for a in sys.argv[1:]] or list(range(8))
gen = default_solver()   # the workload is drawn from the tight closed loops whatever the measured tolerance is
s = default_solver(solver_opts={'b200': {'tol': float(os.environ.get('AB_TOL', '1e-5'))}})
rows = []
for sh in shards:
    x0, p = batches.make_batch(gen, ("exp1", "exp2"), sh * B, B, bound_scale=True)
    xd, pd = torch.from_numpy(x0).cuda(), torch.from_numpy(p).cuda()
    out = s.solve_batch(xd, pd); torch.cuda.synchronize()
    def timeit(xa, pa, o):
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.solve_batch(xa, pa, o); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    t = timeit(xd, pd, out)
    it = out["iters"].cpu().numpy().copy(); st = out["status"].cpu().numpy().copy()
    order = torch.from_numpy(np.argsort(-it, kind="stable").copy()).cuda()
    xs, ps = xd[order].contiguous(), pd[order].contiguous()
    o2 = s.solve_batch(xs, ps); torch.cuda.synchronize()
    t_sorted = timeit(xs, ps, o2)
    row = dict(shard=sh, ms=t, ms_sorted=t_sorted, ratio=t / t_sorted, iters_mean=float(it.mean()), iters_max=int(it.max()),
               iters_p99=float(np.percentile(it, 99)), fails=int((st != 0).sum()), status={int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))},
               hist=np.bincount(np.minimum(it, 60), minlength=61).tolist())
    rows.append(row)
    print(json.dumps({k: v for k, v in row.items() if k != "hist"}), flush=True)
    idx = np.argsort(-it)[:8]
    np.savez(f"gpurun_out/sweep_long_{sh}.npz", idx=idx, x0=x0[idx], p=p[idx], iters=it[idx], status=st[idx])
json.dump(rows, open("gpurun_out/shard_sweep.json", "w"))
