"""Synthetic OCP instance batches of BASELINE.json's configs.

An instance is one solver call `(x0, p)` as the reference's `BoundMPC.step` would issue it
(BoundMPC.py:310-453).  Batches are built by (1) running the nominal headless closed loop of a
scenario (bound_mpc_node.py:292-372: step -> integrate_joint -> next state) with the CUDA solver,
snapshotting the controller state at every step, and (2) for instance i restoring the snapshot of
step `i mod T`, perturbing the joint state with `default_rng(20261017 + i)` and calling the
host-side `prepare()`.  Odd instances use the cold start, even ones the shifted warm start.
Generated batches are cached as .npz (inputs are never part of a timed region).
"""
import copy
import hashlib
import os
import numpy as np

from . import scenarios
from .bound_mpc import BoundMPC, integrate_joint
from .robot_model import RobotModel, Q_LIM_LOWER, Q_LIM_UPPER, DQ_LIM_LOWER, DQ_LIM_UPPER

SEED0 = 20261017
CACHE_DIR = os.environ.get("BMPC_CACHE", "/tmp/boundmpc_b200_cache")


class Params:
    """The fields of the MPCParams service request the hot path reads (MPCParams.srv:2-11)."""

    def __init__(self, scn, real_time=True):
        self.n, self.nr_segs, self.dt = scn['n'], scn['nr_segs'], scn['dt']
        self.weights = list(scn['weights'])
        self.build = True
        self.real_time = real_time
        self.simulate = self.experiment = self.use_acados = self.learning_based = False


def make_mpc(scn, solver, real_time=True):
    s = copy.deepcopy(scn)
    return BoundMPC(s['p_via'], s['r_via'], [s['p_lower'], s['p_upper']], [s['r_lower'], s['r_upper']], s['bp1'],
                    s['br1'], s['s'], s['e_p_min'], s['e_r_min'], s['e_p_max'], s['e_r_max'], p0=s['p0fk'],
                    params=Params(s, real_time), solver=solver)


def _snapshot(mpc):
    d = {k: v for k, v in mpc.__dict__.items() if k != 'solver'}
    return copy.deepcopy(d)


def _restore(mpc, snap):
    solver = mpc.solver
    mpc.__dict__.update(copy.deepcopy(snap))
    mpc.solver = solver


def nominal_sequence(scn, solver, max_steps=600, record=None):
    """Headless closed loop of one scenario.  Returns the list of per-step
    (controller snapshot, robot state) pairs and per-step solver statistics."""
    mpc = make_mpc(scn, solver)
    rm = RobotModel()
    q, dq, ddq, jerk, v = scn['q0'].copy(), np.zeros(7), np.zeros(7), np.zeros(7), np.zeros(6)
    x_phi_d = np.array([mpc.phi_max[0], 0.0, 0.0])
    snaps, stats = [], []
    for step in range(max_steps):
        p_lie = rm.fk(q)
        snaps.append((_snapshot(mpc), dict(q=q.copy(), dq=dq.copy(), ddq=ddq.copy(), jerk=jerk.copy(), v=v.copy())))
        traj, _, _, t_solve, iters = mpc.step(q, dq, ddq, p_lie, v, x_phi_d, jerk)
        if traj is None:
            raise RuntimeError("nominal sequence: the controller gave up")
        stats.append((iters, t_solve, mpc.solver.stats()['success']))
        if record is not None:
            record.append(traj)
        jm = np.concatenate((jerk[:, None], traj['dddq'][:, :2]), axis=1)
        q, dq, ddq, p_lie, v, _, _ = integrate_joint(rm, jm, q, dq, ddq, mpc.dt)
        jerk = traj['dddq'][:, 0].copy()
        if mpc.phi_max[0] - mpc.phi_current[0] <= 0.01:      # experiment1_runner.py:109
            break
    return snaps, stats, x_phi_d


def perturbed_instance(mpc, snaps, x_phi_d, i, bound_scale=False):
    """Instance i of a batch."""
    rng = np.random.default_rng(SEED0 + i)
    snap, st = snaps[i % len(snaps)]
    _restore(mpc, snap)
    q = st['q'] + rng.normal(0.0, 0.02, 7)
    dq = st['dq'] + rng.normal(0.0, 0.02, 7)
    ddq = st['ddq'] + rng.normal(0.0, 0.05, 7)
    q = np.clip(q, Q_LIM_LOWER + 0.05, Q_LIM_UPPER - 0.05)
    dq = np.clip(dq, DQ_LIM_LOWER + 0.05, DQ_LIM_UPPER - 0.05)
    if bound_scale:
        f = rng.uniform(0.75, 1.25, 4)
        rp = mpc.ref_path
        rp.e_p_min = [v * f[0] for v in rp.e_p_min]
        rp.e_p_max = [v * f[1] for v in rp.e_p_max]
        rp.e_r_min = [v * f[2] for v in rp.e_r_min]
        rp.e_r_max = [v * f[3] for v in rp.e_r_max]
    if i % 2 == 1:
        mpc.prev_solution = None
    rm = mpc.robot_model
    p0 = rm.fk(q)
    v0 = rm.jacobian_fk(q) @ dq
    w0, p, _ = mpc.prepare(q, dq, ddq, p0, v0, x_phi_d, st['jerk'])
    return w0, p


def _cache_path(key):
    os.makedirs(CACHE_DIR, exist_ok=True)
    return os.path.join(CACHE_DIR, hashlib.sha1(key.encode()).hexdigest()[:16] + ".npz")


def make_batch(solver, scenario_names, first, count, n=10, tight=False, bound_scale=False, cache=True):
    """Instances `first .. first+count-1`; instance i uses scenario_names[i % len(scenario_names)]."""
    key = f"v3|{scenario_names}|{first}|{count}|{n}|{tight}|{bound_scale}"
    path = _cache_path(key)
    if cache and os.path.exists(path):
        z = np.load(path)
        return z['x0'], z['p']
    seqs = {}
    for name in set(scenario_names):
        scn = scenarios.experiment1(n=n, tight=tight) if name == 'exp1' else scenarios.experiment2(n=n)
        snaps, stats, xd = nominal_sequence(scn, solver)
        seqs[name] = (make_mpc(scn, solver), snaps, xd)
    x0 = np.empty((count, 44 * n))
    p = np.empty((count, solver.np))
    for j in range(count):
        i = first + j
        mpc, snaps, xd = seqs[scenario_names[i % len(scenario_names)]]
        x0[j], p[j] = perturbed_instance(mpc, snaps, xd, i, bound_scale)
    if cache:
        np.savez(path, x0=x0, p=p)
    return x0, p


# BASELINE.json configs -> generator arguments
CONFIGS = {
    "exp1_1024": dict(scenario_names=("exp1",), count=1024, n=10),
    "exp2_8192": dict(scenario_names=("exp2",), count=8192, n=10),
    "exp1_N20_tight_8192": dict(scenario_names=("exp1",), count=8192, n=20, tight=True),
    "mixed_65536": dict(scenario_names=("exp1", "exp2"), count=65536, n=10, bound_scale=True),
}
