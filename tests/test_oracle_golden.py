"""The CPU oracle against values / derivatives / converged points obtained by executing the
reference's own Python (tests/golden/make_golden.py) and against SURVEY App. D.4."""
import numpy as np
import pytest
from oracle import oracle as O
from tests.util import load, active_set, rel_q_error


@pytest.mark.parametrize("scn", ["exp1", "exp2"])
def test_values_and_first_derivatives(scn):
    G = load(f"nlp_{scn}.npz")
    lbx, ubx, lbg, ubg = O.bounds()
    assert np.array_equal(lbx, G["lbx"]) and np.array_equal(ubx, G["ubx"])
    assert np.array_equal(lbg, G["lbg"]) and np.array_equal(ubg, G["ubg"])
    for i in range(len(G["x"])):
        f, g = O.eval_fg(G["x"][i], G["p"][i])
        assert abs(f - G["f"][i]) <= 1e-13 * abs(G["f"][i])
        assert (np.abs(g - G["g"][i]) <= 1e-12 * np.maximum(1.0, np.abs(G["g"][i]))).all()
        grad, jac, _ = O.derivs(G["x"][i], G["p"][i], np.zeros(430))
        assert np.abs(grad - G["grad"][i]).max() <= 1e-11 * max(1.0, np.abs(G["grad"][i]).max())
        assert np.abs(jac - G["jac"][i]).max() <= 1e-9 * max(1.0, np.abs(G["jac"][i]).max())


@pytest.mark.parametrize("scn", ["exp1", "exp2"])
def test_lagrangian_hessian(scn):
    G = load(f"nlp_{scn}.npz")
    i = int(G["hess_point"])
    _, _, hess = O.derivs(G["x"][i], G["p"][i], G["hess_lam"])
    # golden Hessian = central differences of the reference-executed complex-step gradient
    assert np.abs(hess - G["hess"]).max() <= 2e-6 * np.abs(G["hess"]).max()
    assert np.abs(hess - hess.T).max() <= 1e-12 * np.abs(hess).max()


@pytest.mark.parametrize("scn", ["exp1", "exp2"])
def test_converged_points_of_the_sequence(scn):
    S = load(f"seq_{scn}.npz")
    assert S["ref_kkt_residual"][:, 0].max() < 1e-7      # certified with reference-executed derivatives
    for i in range(0, len(S["step"]), 3):
        r = O.solve(S["x0"][i], S["p"][i], tol=1e-10)
        assert r["status"] == 0
        # (the fixtures were solved with mu_init = 0.1; a different barrier path ends within ~mu_min of the same point)
        assert rel_q_error(r["x"], S["x"][i]) < 1e-8
        assert abs(r["f"] - S["f"][i]) < 1e-9 * abs(S["f"][i])


def test_known_answers_survey_d4():
    S = load("seq_exp1.npz")
    assert int(S["step"][0]) == 0
    r = O.solve(S["x0"][0], S["p"][0], tol=1e-10)
    assert abs(r["f"] - 1963.4557512307) < 1e-8
    u0 = [3.02647252, -15.77695511, -5.7643976, -23.47391524, 12.10841625, -7.609702, -5.92469403, 4.71201605]
    assert np.abs(r["x"][:8] - u0).max() < 1e-6
    rows, bnds = active_set(r)
    assert rows == {(4, 39), (5, 40), (6, 40), (9, 37)}
    assert bnds == {(7, 18, -1)}
    lg = r["lam_g"].reshape(10, 43)
    assert abs(lg[4, 39] - 2.367) < 2e-3 and abs(lg[5, 40] - 17.947) < 2e-3 and abs(lg[6, 40] - 40.198) < 2e-3
    S2 = load("seq_exp2.npz")
    r2 = O.solve(S2["x0"][0], S2["p"][0], tol=1e-10)
    assert abs(r2["f"] - 206.08995711648) < 1e-8
    assert active_set(r2)[0] == set()
