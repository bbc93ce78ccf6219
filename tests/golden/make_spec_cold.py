"""Fixture of the literal SURVEY 8d generator (`batches.make_batch(..., spec=True)`): instances 0..63 of the mixed workload
generated with the host build of the kernel sources as the solver (no GPU needed), of which eight are kept: four
cold-started in the middle of a path (odd indices) that converge with the cold-start repair, two warm-started, two
locally infeasible.     python tests/golden/make_spec_cold.py  ->  tests/golden/spec_cold.npz"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from boundmpc_b200 import batches
from tests.emu.emu_solver import EmuSolver
from oracle import oracle as O

if __name__ == "__main__":
    x0, p = batches.make_batch(EmuSolver(), ("exp1", "exp2"), 0, 64, bound_scale=True, spec=True, cache=False, workers=1)
    st = np.array([O.solve(x0[i], p[i], tol=1e-9)["status"] for i in range(64)])
    odd_ok = [i for i in range(1, 64, 2) if st[i] == 0][:4]
    even_ok = [i for i in range(0, 64, 2) if st[i] == 0][:2]
    bad = [i for i in range(64) if st[i] == 5][:2]
    idx = np.array(odd_ok + even_ok + bad)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "spec_cold.npz"), idx=idx, x0=x0[idx], p=p[idx])
