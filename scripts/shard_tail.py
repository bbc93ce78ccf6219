"""Long-running instances of one shard of the mixed workload (development aid): python scripts/shard_tail.py <shard>."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches
shard = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = 8192
s = default_solver()
x0, p = batches.make_batch(s, ("exp1", "exp2"), shard * B, B, bound_scale=True)
xd, pd = torch.from_numpy(x0).cuda(), torch.from_numpy(p).cuda()
out = s.solve_batch(xd, pd); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = s.solve_batch(xd, pd, out); e1.record(); torch.cuda.synchronize()
it = out["iters"].cpu().numpy(); st = out["status"].cpu().numpy(); kkt = out["kkt"].cpu().numpy()
print(f"shard {shard}: {e0.elapsed_time(e1):.2f} ms, iters mean {it.mean():.2f} max {it.max()}, status counts {dict(zip(*np.unique(st, return_counts=True)))}")
idx = np.argsort(-it)[:12]
print("longest:", [(int(i), int(it[i]), int(st[i]), float(f"{kkt[i]:.1e}")) for i in idx])
np.savez("gpurun_out/shard_long.npz", idx=idx, x0=x0[idx], p=p[idx], iters=it[idx], status=st[idx])
