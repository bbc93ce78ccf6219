"""Closed-loop roll-out of a batch of controllers on the device.

One MPC step of the reference's headless loop (bound_mpc_node.py:292-372: `BoundMPC.step` ->
`integrate_joint` -> next measured state) for B independent controller / robot pairs is three kernel
launches on one stream — `prepare_batch` (parameters + warm start), `solve_batch`, `finish_batch`
(accept / fallback, post-processing, robot advance) — with all state resident in HBM: controller states
[B, 76], window positions, previous solutions, error counts.  The host only enqueues.
"""
import numpy as np


def initial_state(mpc, q0, bound_scale=(1.0, 1.0, 1.0, 1.0)):
    """State vector of a controller at rest at joint position q0 (the start of experiment*_runner.py)."""
    rm = mpc.robot_model
    z7, z6 = np.zeros(7), np.zeros(6)
    x_phi_d = np.array([mpc.phi_max[0], 0.0, 0.0])
    st, sector, prev = mpc.builder_state(np.asarray(q0, float), z7, z7, rm.fk(q0), z6, x_phi_d, z7, bound_scale=bound_scale)
    return st, sector, prev


def rollout(solver, tables, path_id, state, sector, steps, prev=None, record=True):
    """Run `steps` closed-loop MPC steps for the B controllers described by torch CUDA tensors `state` [B, 76],
    `sector` [B] int32, `path_id` [B] int32 and the path `tables` [P, J, 41].  Returns the final (state, sector, prev,
    error_count) and, with record=True, the per-step logs q [steps, B, 7], phi [steps, B], iters, status [steps, B]."""
    import torch
    B, dev = state.shape[0], state.device
    state = state.clone()
    sector = sector.clone()
    prev = torch.zeros((B, solver.n), dtype=torch.float64, device=dev) if prev is None else prev.clone()
    ec = torch.zeros(B, dtype=torch.int32, device=dev)
    nxt = torch.empty_like(state)
    bo = so = fo = None
    log = {"q": [], "phi": [], "iters": [], "status": [], "error_count": []}
    for _ in range(steps):
        bo = solver.prepare_batch(tables, path_id, sector, state, prev, bo)
        so = solver.solve_batch(bo["x0"], bo["p"], so)
        fo = solver.finish_batch(tables, path_id, sector, state, so, prev, ec, True, {"traj": fo["traj"], "state": nxt} if fo else {"state": nxt})
        state, nxt = fo["state"], state
        if record:
            log["q"].append(state[:, 0:7].clone()); log["phi"].append(state[:, 40].clone())
            log["iters"].append(so["iters"].clone()); log["status"].append(so["status"].clone()); log["error_count"].append(ec.clone())
    out = {"state": state, "sector": sector, "prev": prev, "error_count": ec}
    if record:
        out.update({k: torch.stack(v) for k, v in log.items()})
    return out
