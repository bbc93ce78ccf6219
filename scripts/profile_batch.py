"""One batched solve bracketed by cudaProfilerStart/Stop for `ncu --profile-from-start off` (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
s = default_solver()
x0, p = batches.make_batch(s, ("exp1", "exp2"), 0, B, bound_scale=True)
xd, pd = torch.from_numpy(x0).cuda(), torch.from_numpy(p).cuda()
out = s.solve_batch(xd, pd); torch.cuda.synchronize()
torch.cuda.profiler.start()
s.solve_batch(xd, pd, out); torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", int((out["status"] == 0).sum()), "of", B)
