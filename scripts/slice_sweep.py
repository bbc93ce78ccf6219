"""Pass-A length of the two-pass scheduling (development aid): kernel time of one shard for BMPC_SLICE_ITERS = 3..8."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches
B = 8192
shard = int(sys.argv[1]) if len(sys.argv) > 1 else 0
s0 = default_solver(solver_opts={'b200': {'tol': float(os.environ.get('AB_TOL', '1e-5'))}})
x0, p = batches.make_batch(default_solver(), ("exp1", "exp2"), shard * B, B, bound_scale=True)
xd, pd = torch.from_numpy(x0).cuda(), torch.from_numpy(p).cuda()
ref = None
for k in [int(a) for a in sys.argv[2:]] or [3, 4, 5, 6, 7, 8]:
    os.environ["BMPC_SLICE_ITERS"] = str(k)
    s = default_solver(solver_opts={'b200': {'tol': float(os.environ.get('AB_TOL', '1e-5'))}})
    out = s.solve_batch(xd, pd); torch.cuda.synchronize()
    best = 1e9
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve_batch(xd, pd, out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    x = out["x"].cpu().numpy()
    same = True if ref is None else bool(np.array_equal(x, ref))
    ref = x if ref is None else ref
    print(f"slice_iters {k}: {best:.2f} ms  {B / best * 1e3:.0f} solves/s  bitwise same as first: {same}", flush=True)
