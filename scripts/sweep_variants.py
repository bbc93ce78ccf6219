"""Device throughput of k_solve for every compiled launch shape on the bench workload (development aid)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
base = default_solver()
x0, p = batches.make_batch(base, ("exp1", "exp2"), 0, B, bound_scale=True)
xd, pd = torch.from_numpy(x0).cuda(), torch.from_numpy(p).cuda()
for threads, ctas in [(128, 3), (160, 3), (192, 3), (256, 2), (384, 1)]:
    os.environ["BMPC_THREADS"], os.environ["BMPC_CTAS_PER_SM"] = str(threads), str(ctas)
    try:
        s = default_solver()
    except Exception as e:
        print(threads, ctas, "unavailable:", e); continue
    out = s.solve_batch(xd, pd); torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve_batch(xd, pd, out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(json.dumps({"threads": threads, "ctas_per_sm": ctas, "B": B, "ms": best, "solves_per_s": B / best * 1e3,
                      "shape": s.launch_shape(), "ok": int((out["status"] == 0).sum()), "iters_mean": float(out["iters"].double().mean())}), flush=True)
