"""Utilisation profile of one launch (development aid; BMPC_LIB = build with -DBMPC_TRACE): busy CTA slots per
millisecond from the per-instance time stamps of the scheduler."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches, _cabi
B = 8192
shard = int(sys.argv[1]) if len(sys.argv) > 1 else 0
gen = default_solver()   # the workload is drawn from the tight closed loops whatever the measured tolerance is
s = default_solver(solver_opts={'b200': {'tol': float(os.environ.get('AB_TOL', '1e-5'))}})
x0, p = batches.make_batch(gen, ("exp1", "exp2"), shard * B, B, bound_scale=True)
xd, pd = torch.from_numpy(x0).cuda(), torch.from_numpy(p).cuda()
out = s.solve_batch(xd, pd); torch.cuda.synchronize()
out = s.solve_batch(xd, pd, out); torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (4 * B))()
_cabi.lib().bmpc_trace(buf, B)
t = np.frombuffer(buf, dtype=np.uint64).reshape(B, 4).astype(np.int64)
t[:, 1] &= ~1; t[:, 3] &= ~1
t0 = t[:, 0].min()
ms = (t - t0) / 1e6
it = out["iters"].cpu().numpy()
resumed = t[:, 2] > 0
span = max(ms[:, 1].max(), ms[resumed, 3].max())
print(f"span {span:.2f} ms, pass A handed out by {ms[:, 0].max():.2f} ms, last slice ends {ms[:, 1].max():.2f} ms, resumed {int(resumed.sum())}")
edges = np.arange(0.0, span + 1.0, 1.0)
busy = np.zeros(len(edges) - 1)
for a, b in [(ms[:, 0], ms[:, 1]), (ms[resumed, 2], ms[resumed, 3])]:
    for k in range(len(busy)):
        busy[k] += np.clip(np.minimum(b, edges[k + 1]) - np.maximum(a, edges[k]), 0, None).sum()
print("busy CTA slots per ms bin:", " ".join(f"{v:.0f}" for v in busy))
tot = busy.sum()
print(f"total busy {tot:.0f} slot-ms = {tot / 444:.2f} ms of a full machine; launch span {span:.2f} ms; utilisation {tot / 444 / span:.3f}")
slice_ms = (ms[:, 1] - ms[:, 0]); res_ms = (ms[resumed, 3] - ms[resumed, 2])
print(f"mean slice {slice_ms.mean() * 1e3:.0f} us, mean resumed run {res_ms.mean() * 1e3:.0f} us; per iteration (all work / all iterations) {tot / it.sum() * 1e3:.1f} us")
last = np.argsort(-np.where(resumed, ms[:, 3], ms[:, 1]))[:10]
print("last finishers (instance, iters, resumed at, end):", [(int(i), int(it[i]), round(float(ms[i, 2]), 2), round(float(max(ms[i, 3], ms[i, 1])), 2)) for i in last])
