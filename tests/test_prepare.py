"""Batched parameter builder (csrc/bmpc_prepare.cuh) against the host mirror of the pre-solve half
of `BoundMPC.step` (boundmpc_b200/bound_mpc.py `prepare`, itself checked against the reference class in
tests/test_host_mirror.py).  CPU: the kernel source compiled for the host (tests/emu); GPU: the C ABI."""
import numpy as np
import pytest
from tests.emu import emu
from tests.emu.emu_solver import EmuSolver
from boundmpc_b200 import batches


def _rel(a, b):
    return float((np.abs(a - b) / np.maximum(1.0, np.abs(b))).max())


@pytest.fixture(scope="module")
def builder_batch():
    return batches.make_builder_batch(EmuSolver(), ("exp1", "exp2"), 0, 256, bound_scale=True)


def test_builder_matches_host_mirror(builder_batch):
    D = builder_batch
    x0, p, sec = emu.prepare(D["tables"], D["path_id"], D["sector"], D["state"], D["prev"])
    assert np.array_equal(sec, D["sector_out"])
    assert np.array_equal(x0, D["x0"])                       # copies and shifts: exact
    assert _rel(p, D["p"]) < 1e-12                           # fp64 arithmetic in a different operation order
    assert (D["sector_out"] > 0).any() and (D["state"][:, 73] == 0).any() and (D["state"][:, 73] != 0).any()


def test_builder_window_slides_over_several_segments(builder_batch):
    """An instance whose stored window position lags behind its path parameter: the builder advances it like
    ReferencePath.update does (ReferencePath.py:190-212)."""
    D = builder_batch
    j = int(np.argmax(D["sector_out"]))
    assert D["sector_out"][j] >= 2
    x0, p, sec = emu.prepare(D["tables"], D["path_id"][j:j + 1], [0], D["state"][j:j + 1], D["prev"][j:j + 1])
    assert sec[0] == D["sector_out"][j] and _rel(p[0], D["p"][j]) < 1e-12


def test_builder_reversing_integrated_omega(builder_batch):
    """BoundMPC.py:325-333: a jump of the measured orientation vector by more than 1.5 rewrites p_rot of the warm start."""
    D = builder_batch
    j = int(np.argmax(D["state"][:, 73] != 0))
    st, prev = D["state"][j].copy(), D["prev"][j].copy()
    st[24:27] += np.array([1.2, -0.9, 0.8])                  # p0 rotation part far from the previous solution
    x0, _, _ = emu.prepare(D["tables"], D["path_id"][j:j + 1], D["sector"][j:j + 1], st, prev)
    w = prev.reshape(10, 44).copy()
    first = w[0, 32:35].copy()
    assert np.linalg.norm(st[24:27] - first) > 1.5
    w[:-1, 32:35] = st[24:27] + (w[1:, 32:35] - first)
    w[-1, 32:35] = w[-2, 32:35]
    w[:-1] = w[1:].copy()
    assert np.array_equal(x0[0], w.ravel())


@pytest.mark.gpu
def test_gpu_builder_matches_host_mirror_and_feeds_the_solver():
    import torch
    from boundmpc_b200.ocp import default_solver
    s = default_solver()
    D = batches.make_builder_batch(s, ("exp1", "exp2"), 0, 1024, bound_scale=True)
    r = s.prepare_batch(D["tables"], D["path_id"], D["sector"], D["state"], D["prev"])          # host-pointer entry
    assert np.array_equal(r["sector"], D["sector_out"])
    assert np.array_equal(r["x0"], D["x0"])
    assert _rel(r["p"], D["p"]) < 1e-12
    # device entry on torch tensors, chained into the solver without leaving the GPU
    dev = torch.device("cuda")
    t = {k: torch.from_numpy(np.ascontiguousarray(D[k])).to(dev) for k in ("tables", "path_id", "sector", "state", "prev")}
    rd = s.prepare_batch(t["tables"], t["path_id"], t["sector"], t["state"], t["prev"])
    assert np.array_equal(rd["p"].cpu().numpy(), r["p"]) and np.array_equal(rd["x0"].cpu().numpy(), r["x0"])
    a = s.solve_batch(rd["x0"], rd["p"])
    b = s.solve_batch(torch.from_numpy(D["x0"]).to(dev), torch.from_numpy(D["p"]).to(dev))
    ok = (a["status"] == 0) & (b["status"] == 0)
    assert int(ok.sum()) >= 800          # (full-size perturbations without the generator's feasibility repair: some instances are infeasible)
    assert bool((a["status"] == b["status"]).all())
    qa, qb = a["x"].cpu().numpy().reshape(-1, 10, 44)[:, :, 8:15], b["x"].cpu().numpy().reshape(-1, 10, 44)[:, :, 8:15]
    okn = ok.cpu().numpy()
    err = np.abs(qa - qb).reshape(len(qa), -1).max(1) / np.abs(qb).reshape(len(qb), -1).max(1)
    assert np.median(err[okn]) < 1e-9 and (err[okn] < 1e-6).mean() > 0.99      # (a few instances have two local solutions)
