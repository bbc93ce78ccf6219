"""How much of a launch is tail?  Times the bench batch as is, with the long-running instances replaced by
ordinary ones, and sorted by iteration count (development aid)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches
B = 8192
s = default_solver()
x0, p = batches.make_batch(s, ("exp1", "exp2"), 0, B, bound_scale=True)
xd, pd = torch.from_numpy(x0).cuda(), torch.from_numpy(p).cuda()
out = s.solve_batch(xd, pd); torch.cuda.synchronize()
it = out["iters"].cpu().numpy().copy()
def timeit(xa, pa, label):
    o = s.solve_batch(xa, pa); torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve_batch(xa, pa, o); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    tot = int(o["iters"].sum())
    print(f"{label}: {best:.2f} ms, {B / best * 1e3:.0f} solves/s, total iterations {tot}, {best * 1e3 / tot * 444:.3f} us per iteration-slot")
timeit(xd, pd, "as is")
for cap in (40, 25, 20):
    idx = np.arange(B); idx[it > cap] = 0
    ii = torch.from_numpy(idx).cuda()
    timeit(xd[ii].contiguous(), pd[ii].contiguous(), f"instances with > {cap} iterations replaced ({int((it > cap).sum())})")
order = torch.from_numpy(np.argsort(-it, kind="stable").copy()).cuda()
timeit(xd[order].contiguous(), pd[order].contiguous(), "sorted, longest first")
