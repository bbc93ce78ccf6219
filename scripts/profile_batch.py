"""One batched solve bracketed by cudaProfilerStart/Stop for `ncu --profile-from-start off` (development aid).
usage: profile_batch.py [B] [config]   config = a key of batches.CONFIGS (default mixed_65536)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
name = sys.argv[2] if len(sys.argv) > 2 else "mixed_65536"
c = dict(batches.CONFIGS[name]); c.pop("count"); n = c.pop("n")
gen = default_solver(N=n)
s = default_solver(N=n, solver_opts={'b200': {'tol': float(os.environ.get('AB_TOL', '1e-5'))}})
x0, p = batches.make_batch(gen, first=0, count=B, n=n, **c)
xd, pd = torch.from_numpy(x0).cuda(), torch.from_numpy(p).cuda()
out = s.solve_batch(xd, pd); torch.cuda.synchronize()
torch.cuda.profiler.start()
s.solve_batch(xd, pd, out); torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", int((out["status"] == 0).sum()), "of", B, "iters mean", float(out["iters"].double().mean()), "launch shape", s.launch_shape())
