"""Scheduler timeline of one launch (development aid; BMPC_LIB = build with -DBMPC_TRACE)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches, _cabi
B = 8192
s = default_solver()
x0, p = batches.make_batch(s, ("exp1", "exp2"), 0, B, bound_scale=True)
xd, pd = torch.from_numpy(x0).cuda(), torch.from_numpy(p).cuda()
out = s.solve_batch(xd, pd); torch.cuda.synchronize()
out = s.solve_batch(xd, pd, out); torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (4 * B))()
_cabi.lib().bmpc_trace(buf, B)
t = np.frombuffer(buf, dtype=np.uint64).reshape(B, 4).astype(np.int64)
hard = t[:, 1] & 1
t0 = t[:, 0].min()
ms = (t - t0) / 1e6
it = out["iters"].cpu().numpy()
print("launch span %.2f ms; pass A ends %.2f ms; flagged hard %d" % (ms[:, [1, 3]].max(), ms[:, 1].max(), hard.sum()))
for i in np.argsort(-it)[:10]:
    print(f"instance {i}: iters {it[i]} hard {hard[i]} slice {ms[i,0]:.2f}-{ms[i,1]:.2f} resumed {ms[i,2]:.2f}-{ms[i,3]:.2f}")
last = np.argsort(-ms[:, 3])[:8]
print("last finishers:", [(int(i), int(it[i]), int(hard[i]), round(float(ms[i, 2]), 2), round(float(ms[i, 3]), 2)) for i in last])
np.savez("gpurun_out/trace.npz", t=t, iters=it)
# iteration log of the slowest non-flagged instance, solved alone
L = _cabi.lib()
cand = [int(i) for i in np.argsort(-it)[:12] if not hard[i]][:2]
for i in cand:
    xi, pi = xd[i:i + 1].contiguous(), pd[i:i + 1].contiguous()
    L.bmpc_itlog(None, 1)
    o = s.solve_batch(xi, pi); torch.cuda.synchronize()
    L.bmpc_itlog(None, 0)
    lg = (ctypes.c_double * (12 * 500))()
    L.bmpc_itlog(lg, -1)
    lg = np.frombuffer(lg, dtype=np.float64).reshape(500, 12)
    n = int(o["iters"][0])
    print(f"--- instance {i} alone: iters {n} status {int(o['status'][0])}")
    for k in range(min(n, 100)):
        e0, dinf, pinf, mu, al, apr, adu, dw, th, ph, dphi, fl = lg[k]
        print(f"  it {k:3d} e0 {e0:.3e} d {dinf:.2e} p {pinf:.2e} mu {mu:.1e} alpha {al:.3e} (max {apr:.3e}) adu {adu:.2e} dw {dw:.1e} theta {th:.3e} phi {ph:.12e} dphi {dphi:.2e} flags {int(fl)}")
    np.savez(f"gpurun_out/slow_{i}.npz", x0=x0[i], p=p[i])
