"""Whole-path closed loop of a fleet on the device (development aid): 4,096 robots, experiment1 / experiment2 alternating,
per-robot bound widths, until every robot has reached the end of its path (or 220 steps).  Reports convergence and
fallback statistics per step range."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches, scenarios
from boundmpc_b200.rollout import initial_state, rollout
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 220
s = default_solver()
dev = torch.device("cuda")
st, sec, tabs, phimax = [], [], [], []
for nm in ("exp1", "exp2"):
    scn = scenarios.experiment1(n=10) if nm == "exp1" else scenarios.experiment2(n=10)
    m = batches.make_mpc(scn, batches._BoundsOnly(s.bounds()))
    a, b_, _ = initial_state(m, scn['q0'])
    st.append(a); sec.append(b_); tabs.append(m.ref_path.path_table()); phimax.append(m.phi_max[0])
J = max(t.shape[0] for t in tabs)
T = np.zeros((2, J, 41))
for k, t in enumerate(tabs):
    T[k, :len(t)] = t; T[k, len(t):] = t[-1]
pid = (np.arange(B) % 2).astype(np.int32)
state = np.stack([st[k] for k in pid])
state[:, 53:57] = np.random.default_rng(20261017).uniform(1.0, 1.25, (B, 4))
t0 = time.perf_counter()
ro = rollout(s, torch.from_numpy(T).to(dev), torch.from_numpy(pid).to(dev), torch.from_numpy(state).to(dev),
             torch.from_numpy(np.array([sec[k] for k in pid], np.int32)).to(dev), steps, record=True)
torch.cuda.synchronize()
el = time.perf_counter() - t0
status, iters, ec, phi = (ro[k].cpu().numpy() for k in ("status", "iters", "error_count", "phi"))
pm = np.array(phimax)[pid]
done_step = np.array([np.argmax(pm[b] - phi[:, b] <= 0.01) if (pm[b] - phi[:, b] <= 0.01).any() else -1 for b in range(B)])
print(f"{B} robots x {steps} steps in {el:.2f} s = {B * steps / el:.0f} MPC steps/s")
for k, nm in enumerate(("exp1", "exp2")):
    m = pid == k
    d = done_step[m]
    print(f"{nm}: reached the end of the path: {(d >= 0).sum()} of {m.sum()} (steps p50 {np.median(d[d >= 0]) if (d >= 0).any() else -1:.0f}, max {d.max()})")
    act = np.array([[done_step[b] < 0 or t <= done_step[b] for b in np.flatnonzero(m)] for t in range(steps)])
    stt, itt, ect = status[:, m], iters[:, m], ec[:, m]
    print(f"   while under way: solves {act.sum()}, converged {(stt[act] == 0).mean():.4f}, status hist {dict(zip(*np.unique(stt[act], return_counts=True)))}, "
          f"iterations mean {itt[act].mean():.2f} max {itt[act].max()}, steps on a fallback solution {(ect[act] > 0).mean():.4f}, max error_count {ect[act].max()}")
