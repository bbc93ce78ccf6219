"""Per-phase cycle accounting of k_solve (development aid; needs `python -m boundmpc_b200.build --timing`).
usage: BMPC_LIB=boundmpc_b200/libboundmpc_b200_timing.so python scripts/phase_timing.py [B]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches, _cabi

NAMES = {0: "full: integrate+sincos", 1: "full: path<1> | fk -> kin residual, curvature, jacobian rows", 2: "full: end barrier",
         3: "values: integrate+sincos", 4: "values: fk + kin residual | path<0>", 6: "kkt error + reduce",
         7: "kkt_solve init + prefetch", 8: "riccati 1: add_W", 9: "riccati 2a: Y integrator pass, tv", 10: "riccati 2b: Y dmma",
         11: "riccati 3: Qu, qv", 12: "riccati 4: chol+gains | Qss pass + prefetch", 13: "riccati 5: P dmma, pv",
         14: "forward sweep (warp 0) | staging (other warps)", 15: "adjoint staging + rhs", 16: "adjoint sweep", 17: "kkt_prepare", 18: "step parts + ftb reduce",
         19: "ls: trial point", 20: "ls: merit reduce", 21: "accept step", 22: "init point", 23: "report", 30: "  riccati 5: dmma tiles (warp 0)", 31: "  riccati 5: pv", 35: "  full: path blocks", 36: "  full: grad f", 40: "  riccati 5: flag check", 42: "  riccati 4: factorisation + solves (thread 0)", 48: "  [thread 64] up to phase 4", 49: "  [thread 64] riccati 4: prefetch issue", 50: "  [thread 64] riccati 4: Qss pass 1", 51: "  [thread 64] riccati 4: Qss pass 2", 52: "  [thread 64] riccati 4: cp.async wait", 44: "  [thread 64] everything up to phase 5", 45: "  [thread 64] riccati 5 tiles",
         46: "  [thread 64] riccati 5 pv", 47: "  [thread 64] riccati 5 end barrier"}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
gen = default_solver()   # the workload is drawn from the tight closed loops whatever the measured tolerance is
s = default_solver(solver_opts={'b200': {'tol': float(os.environ.get('AB_TOL', '1e-5'))}})
x0, p = batches.make_batch(gen, ("exp1", "exp2"), 0, B, bound_scale=True)
xd, pd = torch.from_numpy(x0).cuda(), torch.from_numpy(p).cuda()
out = s.solve_batch(xd, pd); torch.cuda.synchronize()
L = _cabi.lib()
buf = (ctypes.c_ulonglong * 64)()
L.bmpc_phase_cycles(buf, 1)
out = s.solve_batch(xd, pd, out); torch.cuda.synchronize()
L.bmpc_phase_cycles(buf, 0)
iters = float(out["iters"].double().sum())
tot = sum(buf[i] for i in range(44))
print(f"B={B} iterations={iters:.0f} total thread-0 cycles per iteration: {tot / iters:.0f}")
for i in range(62):
    if buf[i]:
        print(f"{i:3d} {NAMES.get(i, '?'):55s} {buf[i] / iters:10.0f} cyc/iter {100 * buf[i] / tot:5.1f}%")
