"""SM / hardware warp slot of the warps of a few CTAs of k_solve (development aid).
usage: python -m boundmpc_b200.build --out=/tmp/probe.so -DBMPC_PROBE_SLOTS; BMPC_LIB=/tmp/probe.so python scripts/probe_slots.py | sort -u"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches
s = default_solver()
x0, p = batches.make_batch(s, ("exp1", "exp2"), 0, 8192, bound_scale=True)
xd, pd = torch.from_numpy(x0).cuda(), torch.from_numpy(p).cuda()
out = s.solve_batch(xd, pd); torch.cuda.synchronize()
