"""Generate the golden fixtures in tests/golden/ (run HERE, needs /root/reference).

    python tests/golden/make_golden.py

What is recorded, all obtained by EXECUTING the reference's own Python through
tests/golden/refexec (the numeric `casadi` stand-in):
  * nlp_<scn>.npz   — (x, p) points with f, g evaluated by the reference's
    `setup_optimization_problem` code, complex-step grad f / Jac g of that same code, and one
    finite-difference Lagrangian Hessian per scenario;
  * seq_<scn>.npz   — inputs (x0, p) the UNMODIFIED reference `BoundMPC.step` hands to its solver
    at selected steps of the headless closed loop (bound_mpc_node.py:292-372 restated in
    `closed_loop`), together with the converged KKT point of each (oracle, tol 1e-10) and the
    KKT residual of that point measured with the reference-executed derivatives — this is what
    certifies the oracle's solutions as KKT points of the *reference's* NLP.
No Ipopt output exists (casadi is not installable here): the fixtures pin values, derivatives
and converged points, not Ipopt's iterate path.
"""
import os
import sys
import time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from refexec import harness as H  # noqa: E402
from boundmpc_b200 import scenarios  # noqa: E402
from oracle import oracle as O  # noqa: E402


def closed_loop(name, max_steps=400, tol=1e-10, on_step=None):
    scn = scenarios.SCENARIOS[name]()
    recs = []

    def backend(x0, lbx, ubx, lbg, ubg, p):
        x0 = np.array(x0, float)
        p = np.array(p, float)
        r = O.solve(x0, p, tol=tol)
        recs.append(dict(x0=x0, p=p, **r))
        return (dict(x=r['x'], g=r['g'], lam_g=r['lam_g'], lam_x=r['lam_x'], f=r['f']),
                dict(iter_count=r['iters'], success=r['status'] == 0, return_status=str(r['status'])))

    mpc = H.make_reference_mpc(scn, backend)
    rm = H.reference_robot_model()
    integrate_joint = H.reference_integrate_joint()
    q, dq, ddq, jerk, v = scn['q0'].copy(), np.zeros(7), np.zeros(7), np.zeros(7), np.zeros(6)
    x_phi_d = np.array([mpc.phi_max[0], 0, 0])
    states = []
    for step in range(max_steps):
        p_lie, _, _ = rm.forward_kinematics(q, dq)
        states.append(dict(q=q.copy(), dq=dq.copy(), ddq=ddq.copy(), jerk=jerk.copy(), p_lie=p_lie.copy(), v=v.copy(),
                           phi=mpc.phi_current.copy()))
        out = mpc.step(q, dq, ddq, p_lie, v, x_phi_d, jerk)
        assert out[0] is not None
        traj = out[0]
        jm = np.concatenate((jerk[:, None], traj['dddq'][:, :2]), axis=1)
        q, dq, ddq, p_lie, v, a, j = integrate_joint(rm, jm, q, dq, ddq, mpc.dt)
        jerk = traj['dddq'][:, 0].copy()
        if mpc.phi_max[0] - mpc.phi_current[0] <= 0.01:
            break
    return recs, states


def main():
    nlp = H.RefNLP()
    rng = np.random.default_rng(20261017)
    for name in ('exp1', 'exp2'):
        t0 = time.time()
        recs, states = closed_loop(name)
        T = len(recs)
        assert all(r['status'] == 0 for r in recs), "oracle failed on the nominal sequence"
        picks = sorted(set([0, 1, 2, 5, 10, 20, 30, 40, 45, 46, 50, 60, 80, 100, 120, 140, T - 2, T - 1]) & set(range(T)))
        seq = {k: np.array([recs[i][k] for i in picks]) for k in ('x0', 'p', 'x', 'g', 'lam_g', 'lam_x', 'f', 'iters', 'kkt')}
        seq['step'] = np.array(picks)
        seq['n_steps'] = np.array(T)
        seq['iters_all'] = np.array([r['iters'] for r in recs])
        seq['f_all'] = np.array([r['f'] for r in recs])
        # certify: KKT residual of the oracle point with the reference-executed derivatives
        res = []
        for i in picks:
            r = recs[i]
            f0, g0, grad, jac = nlp.eval_derivs(r['x'], r['p'])
            res.append([np.abs(grad + jac.T @ r['lam_g'] + r['lam_x']).max(), abs(f0 - r['f']), np.abs(g0 - r['g']).max()])
        seq['ref_kkt_residual'] = np.array(res)
        print(name, 'steps', T, 'max ref-KKT residual', seq['ref_kkt_residual'][:, 0].max(), 'time %.0fs' % (time.time() - t0))
        assert seq['ref_kkt_residual'][:, 0].max() < 1e-7
        np.savez_compressed(os.path.join(HERE, f'seq_{name}.npz'), **seq)
        # function values / derivatives at the solver inputs and at perturbed points
        pts_x, pts_p, F, G, GR, JAC = [], [], [], [], [], []
        for i in [picks[0], picks[len(picks) // 2], picks[-1]]:
            for pert in (0.0, 0.05):
                x = recs[i]['x0'] + pert * rng.normal(size=nlp.n)
                if pert > 0:
                    x[41::44] = rng.uniform(0, recs[i]['p'][460], nlp.N)   # phi over all window segments
                f0, g0, grad, jac = nlp.eval_derivs(x, recs[i]['p'])
                pts_x.append(x); pts_p.append(recs[i]['p']); F.append(f0); G.append(g0); GR.append(grad); JAC.append(jac)
        lam = rng.normal(size=nlp.m)
        Hs = nlp.lag_hess(pts_x[1], pts_p[1], lam)
        np.savez_compressed(os.path.join(HERE, f'nlp_{name}.npz'), x=np.array(pts_x), p=np.array(pts_p), f=np.array(F),
                            g=np.array(G), grad=np.array(GR), jac=np.array(JAC), hess_lam=lam, hess=Hs, hess_point=np.array(1),
                            lbx=nlp.lbx, ubx=nlp.ubx, lbg=nlp.lbg, ubg=nlp.ubg)
        print(name, 'nlp fixtures written, time %.0fs' % (time.time() - t0))


if __name__ == '__main__':
    main()
