"""Pass-A length x tolerance x shard (development aid): kernel time of every shard of the 65,536-instance workload on one GPU
for BMPC_SLICE_ITERS in SLICES at tol in TOLS.  python scripts/shard_slice.py [shards...]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches
B = 8192
SLICES = [int(a) for a in os.environ.get("SLICES", "4,5,6").split(",")]
TOLS = [float(a) for a in os.environ.get("TOLS", "1e-5,1e-9").split(",")]
shards = [int(a) for a in sys.argv[1:]] or list(range(8))
gen = default_solver()
solvers = {}
for tol in TOLS:
    for k in SLICES:
        os.environ["BMPC_SLICE_ITERS"] = str(k)
        solvers[(tol, k)] = default_solver(solver_opts={"b200": {"tol": tol}})
os.environ.pop("BMPC_SLICE_ITERS")
rows = []
for sh in shards:
    x0, p = batches.make_batch(gen, ("exp1", "exp2"), sh * B, B, bound_scale=True)
    xd, pd = torch.from_numpy(x0).cuda(), torch.from_numpy(p).cuda()
    row = {"shard": sh}
    for (tol, k), s in solvers.items():
        out = s.solve_batch(xd, pd); torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.solve_batch(xd, pd, out); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        row[f"tol{tol:g}_slice{k}_ms"] = round(best, 2)
        row[f"tol{tol:g}_ok"] = int((out["status"] == 0).sum()); row[f"tol{tol:g}_itmax"] = int(out["iters"].max())
    rows.append(row)
    print(json.dumps(row), flush=True)
json.dump(rows, open("gpurun_out/shard_slice.json", "w"))
