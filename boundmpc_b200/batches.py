"""Synthetic OCP instance batches of BASELINE.json's configs.

An instance is one solver call `(x0, p)` as the reference's `BoundMPC.step` would issue it
(BoundMPC.py:310-453).  Batches are built by (1) running the nominal headless closed loop of a
scenario (bound_mpc_node.py:292-372: step -> integrate_joint -> next state) with the CUDA solver,
snapshotting the controller state at every step, and (2) for instance i restoring the snapshot of
step `i mod T`, perturbing the joint state with `default_rng(20261017 + i)` and calling the
host-side `prepare()`.  As in the reference, the cold start (BoundMPC.py:316-321) is used by the instances
drawn from step 0 only (its all-zero path parameter is meaningless further along the path); all
others use the shifted previous solution (BoundMPC.py:373-375).
Generated batches are cached as .npz (inputs are never part of a timed region).

Two generators.  The default ("repaired") one is the bench workload: sensor-noise sized perturbations (sigma_q 5e-3),
bound widths x U(1, 1.25), and a perturbation that puts the start further outside the error bounds than the nominal
closed loop ever is gets halved (up to five times) -- 99.9 % of its instances are solvable.  `spec=True` is SURVEY 8d
taken literally: sigma_q = 0.02 rad, sigma_dq = 0.02, sigma_ddq = 0.05, every odd instance cold-started, bound widths
x U(0.75, 1.25), nothing repaired; a 0.02 rad joint perturbation moves the tool by centimetres while the tight segments
allow millimetres, so a large share of those instances starts outside its bounds with no feasible way back within the
jerk limit (they end as "locally infeasible" or fall back on the previous solution in the closed loop).  bench.py
reports both side by side.
"""
import copy
import hashlib
import os
import time
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


def nominal_sequence(scn, solver, max_steps=600, record=None, device_step=True, real_time=True):
    """Headless closed loop of one scenario.  Returns the list of per-step
    (controller snapshot, robot state) pairs and per-step statistics (iterations, time inside the solver call, success,
    wall time of the whole `BoundMPC.step`).  device_step=False: the numpy mirror around solver(x0=, p=)."""
    mpc = make_mpc(scn, solver, real_time)
    mpc.device_step = device_step
    rm = RobotModel()
    q, dq, ddq, jerk, v = scn['q0'].copy(), np.zeros(7), np.zeros(7), np.zeros(7), np.zeros(6)
    x_phi_d = np.array([mpc.phi_max[0], 0.0, 0.0])
    snaps, stats = [], []
    for step in range(max_steps):
        p_lie = rm.fk(q)
        snaps.append((_snapshot(mpc), dict(q=q.copy(), dq=dq.copy(), ddq=ddq.copy(), jerk=jerk.copy(), v=v.copy())))
        t_w = time.perf_counter()
        traj, _, _, t_solve, iters = mpc.step(q, dq, ddq, p_lie, v, x_phi_d, jerk)
        t_w = time.perf_counter() - t_w
        if traj is None:
            raise RuntimeError("nominal sequence: the controller gave up")
        stats.append((iters, t_solve, mpc.solver.stats()['success'], t_w))
        if record is not None:
            record.append(traj)
        jm = np.concatenate((jerk[:, None], traj['dddq'][:, :2]), axis=1)
        q, dq, ddq, p_lie, v, _, _ = integrate_joint(rm, jm, q, dq, ddq, mpc.dt)
        jerk = traj['dddq'][:, 0].copy()
        if mpc.phi_max[0] - mpc.phi_current[0] <= 0.01:      # experiment1_runner.py:109
            break
    return snaps, stats, x_phi_d


# perturbation of the joint state (standard deviations at scale 1): sensor-noise sized.  The
# error bounds of the scenarios are as tight as 1e-4 m (exp2, middle segment) and the first
# horizon nodes cannot be moved by the jerk input (h^3/24 * 35 = 1.5e-3 rad), so a perturbed
# start can make the NLP infeasible; `make_batch` halves the perturbation of such an instance
# until its zero-jerk roll-out is no further outside the bounds than the unperturbed one.
SIGMA_Q, SIGMA_DQ, SIGMA_DDQ = 5e-3, 2e-2, 5e-2
SPEC_SIGMA = (2e-2, 2e-2, 5e-2)          # q, dq, ddq
SCALES = (1.0, 0.5, 0.25, 0.125, 0.0625, 0.0)
CHECK_STAGES = 3


def perturbed_instance(mpc, snaps, x_phi_d, i, bound_scale=False, scale=1.0, spec=False):
    """Instance i of a batch with the perturbation multiplied by `scale`.
    Returns the warm start, the parameter vector and the zero-jerk roll-out used as
    feasibility probe.  spec: the literal SURVEY 8d generator (module docstring)."""
    rng = np.random.default_rng(SEED0 + i)
    snap, st = snaps[i % len(snaps)]
    _restore(mpc, snap)
    sq, sdq, sddq = SPEC_SIGMA if spec else (SIGMA_Q, SIGMA_DQ, SIGMA_DDQ)
    nq, ndq, nddq = rng.normal(0.0, sq, 7), rng.normal(0.0, sdq, 7), rng.normal(0.0, sddq, 7)
    q = np.clip(st['q'] + scale * nq, Q_LIM_LOWER + 0.05, Q_LIM_UPPER - 0.05)
    dq = np.clip(st['dq'] + scale * ndq, DQ_LIM_LOWER + 0.05, DQ_LIM_UPPER - 0.05)
    ddq = st['ddq'] + scale * nddq
    if bound_scale:
        # repaired generator: widths x U(1, 1.25) -- enlarging e_min / e_max enlarges the quartic bound everywhere, so
        # the nominal closed-loop state stays feasible (shrinking them would not); spec: x U(0.75, 1.25)
        f = rng.uniform(0.75 if spec else 1.0, 1.25, 4)
        rp = mpc.ref_path
        rp.e_p_min = [v * f[0] for v in rp.e_p_min]
        rp.e_p_max = [v * f[1] for v in rp.e_p_max]
        rp.e_r_min = [v * f[2] for v in rp.e_r_min]
        rp.e_r_max = [v * f[3] for v in rp.e_r_max]
    if spec and i % 2 == 1:
        mpc.prev_solution = None         # odd instances: cold start (BoundMPC.py:316-321)
    rm = mpc.robot_model
    p0 = rm.fk(q)
    v0 = rm.jacobian_fk(q) @ dq
    w0, p, _ = mpc.prepare(q, dq, ddq, p0, v0, x_phi_d, st['jerk'])
    return w0, p, _rollout(rm, w0, p, mpc.dt, CHECK_STAGES)


def builder_instance(mpc, snaps, x_phi_d, i, bound_scale=False, scale=1.0):
    """Instance i as input of the CUDA parameter builder (`BatchSolver.prepare_batch`): state vector, window position and
    previous solution, next to the (x0, p) the host mirror's `prepare()` builds from the same controller state."""
    rng = np.random.default_rng(SEED0 + i)
    snap, st = snaps[i % len(snaps)]
    _restore(mpc, snap)
    nq, ndq, nddq = rng.normal(0.0, SIGMA_Q, 7), rng.normal(0.0, SIGMA_DQ, 7), rng.normal(0.0, SIGMA_DDQ, 7)
    q = np.clip(st['q'] + scale * nq, Q_LIM_LOWER + 0.05, Q_LIM_UPPER - 0.05)
    dq = np.clip(st['dq'] + scale * ndq, DQ_LIM_LOWER + 0.05, DQ_LIM_UPPER - 0.05)
    ddq = st['ddq'] + scale * nddq
    f = rng.uniform(1.0, 1.25, 4) if bound_scale else np.ones(4)
    rm = mpc.robot_model
    p0 = rm.fk(q)
    v0 = rm.jacobian_fk(q) @ dq
    state, sector, prev = mpc.builder_state(q, dq, ddq, p0, v0, x_phi_d, st['jerk'], bound_scale=f)
    if bound_scale:
        rp = mpc.ref_path
        rp.e_p_min = [v * f[0] for v in rp.e_p_min]
        rp.e_r_min = [v * f[1] for v in rp.e_r_min]
        rp.e_p_max = [v * f[2] for v in rp.e_p_max]
        rp.e_r_max = [v * f[3] for v in rp.e_r_max]
    w0, p, _ = mpc.prepare(q, dq, ddq, p0, v0, x_phi_d, st['jerk'])
    return state, sector, prev, w0, p, int(mpc.ref_path.sector)


def make_builder_batch(solver, scenario_names, first, count, n=10, bound_scale=False):
    """Inputs of the parameter builder for instances `first .. first+count-1` plus the host mirror's outputs.
    Returns dict(tables [P, J, 38], path_id, sector, state, prev, x0, p, sector_out)."""
    names = sorted(set(scenario_names))
    seqs, tabs = {}, []
    for name in names:
        scn = scenarios.experiment1(n=n) if name == 'exp1' else scenarios.experiment2(n=n)
        snaps, _, xd = nominal_sequence(scn, solver)
        mpc = make_mpc(scn, _BoundsOnly(solver.bounds()))
        seqs[name] = (mpc, snaps, xd)
        tabs.append(mpc.ref_path.path_table())
    J = max(t.shape[0] for t in tabs)
    tables = np.zeros((len(tabs), J, tabs[0].shape[1]))
    for k, t in enumerate(tabs):
        tables[k, :t.shape[0]] = t
        tables[k, t.shape[0]:] = t[-1]
    out = dict(tables=tables, path_id=np.empty(count, np.int32), sector=np.empty(count, np.int32), sector_out=np.empty(count, np.int32),
               state=np.empty((count, BoundMPC.PS_SIZE)), prev=np.empty((count, 44 * n)), x0=np.empty((count, 44 * n)),
               p=np.empty((count, solver.np)))
    for j in range(count):
        i = first + j
        name = scenario_names[i % len(scenario_names)]
        mpc, snaps, xd = seqs[name]
        st, sec, prev, w0, p, sec_out = builder_instance(mpc, snaps, xd, i, bound_scale)
        out['path_id'][j], out['sector'][j], out['sector_out'][j] = names.index(name), sec, sec_out
        out['state'][j], out['prev'][j], out['x0'][j], out['p'][j] = st, prev, w0, p
    return out


def _rollout(rm, w0, p, h, stages):
    """w0 with its first `stages` nodes replaced by the zero-jerk continuation of the initial
    state in p (dynamics of SURVEY App. A.4): every equality row of those nodes is zero, so
    their inequality rows say whether the start itself is inside the error bounds."""
    x = np.array(w0, float).reshape(-1, 44).copy()
    q, dq, ddq = p[0:7].copy(), p[7:14].copy(), p[14:21].copy()
    phi, dphi, ddphi = p[21:24]
    um = p[81:89].copy()
    prot = p[27:30].copy()
    for k in range(stages):
        om0 = np.ravel(rm.omega_ee(q, dq))
        qn = q + h * dq + h * h / 2 * ddq + h ** 3 / 8 * um[:7]
        dqn = dq + h * ddq + h * h / 3 * um[:7]
        ddqn = ddq + h / 2 * um[:7]
        phin = phi + h * dphi + h * h / 2 * ddphi + h ** 3 / 8 * um[7]
        dphin = dphi + h * ddphi + h * h / 3 * um[7]
        ddphin = ddphi + h / 2 * um[7]
        om1 = np.ravel(rm.omega_ee(qn, dqn))
        prot = prot + h / 2 * (om0 + om1)
        x[k, 0:8] = 0.0
        x[k, 8:15], x[k, 15:22], x[k, 22:29] = qn, dqn, ddqn
        x[k, 29:32], x[k, 32:35] = np.ravel(rm.fk_pos(qn)), prot
        x[k, 35:38], x[k, 38:41] = np.ravel(rm.velocity_ee(qn, dqn)), om1
        x[k, 41:44] = phin, dphin, ddphin
        q, dq, ddq, phi, dphi, ddphi, um = qn, dqn, ddqn, phin, dphin, ddphin, np.zeros(8)
    return x.ravel()


def bound_excess(d, N, stages=CHECK_STAGES):
    """max over the first nodes and the five interval pairs of (|m| - h) / h from the interval-form
    rows d [B, 12 N] of `eval_batch` (pairs (m - h, -m - h) at columns 2.. of every node)."""
    d = np.asarray(d).reshape(len(d), N, 12)[:, :stages, 2:].reshape(len(d), stages, 5, 2)
    h = -0.5 * (d[..., 0] + d[..., 1])
    m = 0.5 * (d[..., 0] - d[..., 1])
    return ((np.abs(m) - h) / h).reshape(len(d), -1).max(axis=1)


def _cache_path(key):
    os.makedirs(CACHE_DIR, exist_ok=True)
    return os.path.join(CACHE_DIR, hashlib.sha1(key.encode()).hexdigest()[:16] + ".npz")


class _BoundsOnly:
    """What BoundMPC.__init__ needs from a solver handle; lets generator workers build
    controller objects without touching CUDA."""

    def __init__(self, bounds):
        self._b = bounds

    def bounds(self):
        return self._b


_GEN = {}


def _gen_init(seqs, bounds, bound_scale, spec=False):
    _GEN['b'] = bound_scale
    _GEN['spec'] = spec
    _GEN['seqs'] = {name: (make_mpc(scn, _BoundsOnly(bounds)), snaps, xd) for name, (scn, snaps, xd) in seqs.items()}


def _gen_one(job):
    i, name, scale = job
    mpc, snaps, xd = _GEN['seqs'][name]
    return perturbed_instance(mpc, snaps, xd, i, _GEN['b'], scale, _GEN['spec'])


def make_batch(solver, scenario_names, first, count, n=10, tight=False, bound_scale=False, cache=True, workers=None,
               return_scales=False, spec=False):
    """Instances `first .. first+count-1`; instance i uses scenario_names[i % len(scenario_names)].
    `solver` runs the nominal closed loops (any object with the solver call surface) and, for the repaired generator,
    evaluates the bound excess of the candidates (`eval_batch`)."""
    key = f"v7|{scenario_names}|{first}|{count}|{n}|{tight}|{bound_scale}|{SIGMA_Q}|{SIGMA_DQ}|{SIGMA_DDQ}|{spec}"
    path = _cache_path(key)
    if cache and os.path.exists(path):
        z = np.load(path)
        return (z['x0'], z['p'], z['scale']) if return_scales else (z['x0'], z['p'])
    seqs = {}
    for name in sorted(set(scenario_names)):
        scn = scenarios.experiment1(n=n, tight=tight) if name == 'exp1' else scenarios.experiment2(n=n)
        snaps, stats, xd = nominal_sequence(scn, solver)
        seqs[name] = (scn, snaps, xd)
    bounds = solver.bounds()
    ids = np.arange(first, first + count)
    names = [scenario_names[i % len(scenario_names)] for i in ids]
    x0 = np.empty((count, 44 * n))
    p = np.empty((count, solver.np))
    used = np.full(count, -1.0)
    if workers is None:
        workers = min(16, os.cpu_count() or 1)
    pool = None
    if workers > 1 and count >= 64:
        import multiprocessing as mp
        pool = mp.get_context("fork").Pool(workers, initializer=_gen_init, initargs=(seqs, bounds, bound_scale, spec))
        run = lambda jobs: pool.map(_gen_one, jobs, chunksize=max(1, len(jobs) // (workers * 8)))
    else:
        _gen_init(seqs, bounds, bound_scale, spec)
        run = lambda jobs: [_gen_one(j) for j in jobs]
    try:
        def excess(jobs):
            res = run(jobs)
            xr = np.stack([r[2] for r in res])
            pp = np.stack([r[1] for r in res])
            return res, bound_excess(solver.eval_batch(xr, pp, want_jac=False, want_hess=False)["d"], n)
        if spec:                      # literal generator: one pass, nothing repaired
            res = run([(int(i), nm, 1.0) for i, nm in zip(ids, names)])
            for j, r in enumerate(res):
                x0[j], p[j], used[j] = r[0], r[1], 1.0
            pending_done = True
        else:
            pending_done = False
        # unperturbed instances: the level of bound excess the closed loop itself lives with
        _, ex0 = excess([(int(i), nm, 0.0) for i, nm in zip(ids, names)]) if not pending_done else (None, np.zeros(count))
        limit = np.maximum(ex0, -0.02)
        pending = np.arange(count) if not pending_done else np.arange(0)
        for scale in SCALES:
            if len(pending) == 0:
                break
            res, ex = excess([(int(ids[j]), names[j], scale) for j in pending])
            ok = (ex <= limit[pending]) | (scale == 0.0)
            for j, r, o in zip(pending, res, ok):
                if o:
                    x0[j], p[j], used[j] = r[0], r[1], scale
            pending = pending[~ok]
    finally:
        if pool is not None:
            pool.close()
            pool.join()
    if cache:
        np.savez(path, x0=x0, p=p, scale=used)
    return (x0, p, used) if return_scales else (x0, p)


# BASELINE.json configs -> generator arguments
CONFIGS = {
    "exp1_1024": dict(scenario_names=("exp1",), count=1024, n=10),
    "exp2_8192": dict(scenario_names=("exp2",), count=8192, n=10),
    "exp1_N20_tight_8192": dict(scenario_names=("exp1",), count=8192, n=20, tight=True),
    "mixed_65536": dict(scenario_names=("exp1", "exp2"), count=65536, n=10, bound_scale=True),
    "spec_mixed_65536": dict(scenario_names=("exp1", "exp2"), count=65536, n=10, bound_scale=True, spec=True),
}
