"""Solo latency of the longest-running instances of the bench batch (development aid)."""
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
it = out["iters"].cpu().numpy()
order = np.argsort(-it)[:8].tolist() + [0, 1, 2, 3]
for i in order:
    xi, pi = xd[i:i + 1].contiguous(), pd[i:i + 1].contiguous()
    o = s.solve_batch(xi, pi); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o = s.solve_batch(xi, pi, o); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"instance {i}: iters {int(o['iters'][0])} status {int(o['status'][0])} solo {ms:.3f} ms = {ms / max(1, int(o['iters'][0])):.3f} ms/iter")
