"""On-device closed-loop roll-out of many controllers (development aid): B robots x `steps` MPC steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches, scenarios
from boundmpc_b200.rollout import initial_state, rollout
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
s = default_solver()
dev = torch.device("cuda")
tabs, st0, sec0 = [], [], []
for name in ("exp1", "exp2"):
    scn = scenarios.experiment1(n=10) if name == "exp1" else scenarios.experiment2(n=10)
    mpc = batches.make_mpc(scn, batches._BoundsOnly(s.bounds()))
    tabs.append(mpc.ref_path.path_table())
    st, sector, _ = initial_state(mpc, scn['q0'])
    st0.append(st); sec0.append(sector)
J = max(t.shape[0] for t in tabs)
T = np.zeros((2, J, 41))
for k, t in enumerate(tabs):
    T[k, :t.shape[0]] = t; T[k, t.shape[0]:] = t[-1]
rng = np.random.default_rng(1)
pid = (np.arange(B) % 2).astype(np.int32)
state = np.stack([st0[k] for k in pid])
state[:, 53:57] = rng.uniform(1.0, 1.25, (B, 4))            # per-robot bound widths
sector = np.array([sec0[k] for k in pid], np.int32)
args = (s, torch.from_numpy(T).to(dev), torch.from_numpy(pid).to(dev), torch.from_numpy(state).to(dev), torch.from_numpy(sector).to(dev))
rollout(*args, 3, record=False); torch.cuda.synchronize()
t0 = time.perf_counter()
out = rollout(*args, steps, record=True); torch.cuda.synchronize()
el = time.perf_counter() - t0
ok = (out["status"] == 0).float().mean().item()
print(f"B={B} steps={steps}: {el*1e3:.1f} ms wall, {B*steps/el:.0f} MPC steps/s, converged {100*ok:.2f} %, mean iters {out['iters'].float().mean().item():.1f}, "
      f"max error_count {int(out['error_count'].max())}, phi after roll-out: exp1 {out['phi'][-1][0].item():.3f} exp2 {out['phi'][-1][1].item():.3f}")
