"""Quick device timing of the batched solve on replicated golden instances (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from tests.util import load

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
THREADS = int(sys.argv[2]) if len(sys.argv) > 2 else 0
S1, S2 = load("seq_exp1.npz"), load("seq_exp2.npz")
x0 = np.concatenate([S1["x0"], S2["x0"]]); p = np.concatenate([S1["p"], S2["p"]])
idx = np.arange(B) % len(x0)
solver = default_solver(solver_opts={"b200": {"threads": THREADS}})
xd, pd = torch.from_numpy(x0[idx]).cuda(), torch.from_numpy(p[idx]).cuda()
out = solver.solve_batch(xd, pd); torch.cuda.synchronize()
print("status ok", int((out["status"] == 0).sum()), "/", B, "iters mean", float(out["iters"].double().mean()), "max", int(out["iters"].max()))
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = solver.solve_batch(xd, pd, out); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"threads={THREADS} ctas/sm={os.environ.get('BMPC_CTAS_PER_SM','auto')} B={B} {ms:.3f} ms -> {B / ms * 1e3:.0f} solves/s")
t = time.perf_counter(); r = solver.solve_batch(x0[:1], p[:1]); print("B=1 host call %.3f ms" % ((time.perf_counter() - t) * 1e3))
