import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches
s = default_solver()
x0, p = batches.make_batch(s, ("exp1", "exp2"), 0, 8192, bound_scale=True)
xd, pd = torch.from_numpy(x0).cuda(), torch.from_numpy(p).cuda()
out = s.solve_batch(xd, pd); torch.cuda.synchronize()
