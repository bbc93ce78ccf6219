"""Dump the inputs of the longest-running instances of the bench batch (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches
B = 8192
s = default_solver()
x0, p = batches.make_batch(s, ("exp1", "exp2"), 0, B, bound_scale=True)
out = s.solve_batch(x0, p)
it = np.asarray(out["iters"]); st = np.asarray(out["status"])
idx = np.argsort(-it)[:24]
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/hard_instances.npz", idx=idx, x0=x0[idx], p=p[idx], iters=it[idx], status=st[idx], x=np.asarray(out["x"])[idx])
print(list(zip(idx.tolist(), it[idx].tolist(), st[idx].tolist())))
