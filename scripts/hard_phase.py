"""Per-phase cycles of single hard instances (development aid; BMPC_LIB = timing build)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches, _cabi
B = 8192
s = default_solver()
x0, p = batches.make_batch(s, ("exp1", "exp2"), 0, B, bound_scale=True)
xd, pd = torch.from_numpy(x0).cuda(), torch.from_numpy(p).cuda()
L = _cabi.lib()
buf = (ctypes.c_ulonglong * 64)()
for i in [int(a) for a in sys.argv[1:]]:
    xi, pi = xd[i:i + 1].contiguous(), pd[i:i + 1].contiguous()
    o = s.solve_batch(xi, pi); torch.cuda.synchronize()
    L.bmpc_phase_cycles(buf, 1)
    o = s.solve_batch(xi, pi, o); torch.cuda.synchronize()
    L.bmpc_phase_cycles(buf, 0)
    it = int(o["iters"][0])
    tot = sum(buf[k] for k in range(44))
    print(f"instance {i}: iters {it} status {int(o['status'][0])} cycles/iter {tot / it:.0f}")
    print("   per-iteration cycles by phase:", {k: int(buf[k] / it) for k in range(24) if buf[k]})
