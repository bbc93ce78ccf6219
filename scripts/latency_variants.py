"""Single-instance latency (B = 1, device entry, CUDA events) for every compiled launch shape (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from tests.util import load
S1, S2 = load("seq_exp1.npz"), load("seq_exp2.npz")
x0 = np.concatenate([S1["x0"], S2["x0"]]); p = np.concatenate([S1["p"], S2["p"]])
xd, pd = torch.from_numpy(x0).cuda(), torch.from_numpy(p).cuda()
for threads, ctas in [(128, 3), (256, 2), (384, 1), (512, 1)]:
    os.environ["BMPC_THREADS"], os.environ["BMPC_CTAS_PER_SM"] = str(threads), str(ctas)
    s = default_solver()
    ms, its = [], []
    for i in range(len(x0)):
        xi, pi = xd[i:i + 1].contiguous(), pd[i:i + 1].contiguous()
        o = s.solve_batch(xi, pi); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = s.solve_batch(xi, pi, o); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1)); its.append(int(o["iters"][0]))
    ms, its = np.array(ms), np.array(its)
    print(f"threads {threads}: p50 {np.percentile(ms, 50):.3f} ms  mean {ms.mean():.3f} ms  per iteration {ms.sum() / its.sum() * 1e3:.1f} us  (iters mean {its.mean():.1f})")
