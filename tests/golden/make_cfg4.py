"""Fixture of BASELINE configs[3] (experiment1, N = 20, tightened bounds): six instances of the batch
`batches.make_batch(solver, ("exp1",), 0, 48, n=20, tight=True)` generated with the CPU oracle as the solver
(no GPU needed; ~3 minutes).  Three of them converge, three are locally infeasible.
    python tests/golden/make_cfg4.py  ->  tests/golden/cfg4_tight.npz"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from oracle.oracle import OracleSolver
from boundmpc_b200 import batches

if __name__ == "__main__":
    x0, p, sc = batches.make_batch(OracleSolver(N=20), ("exp1",), 0, 48, n=20, tight=True, cache=False, workers=8, return_scales=True)
    idx = np.array([0, 1, 3, 14, 15, 20])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cfg4_tight.npz"), idx=idx, x0=x0[idx], p=p[idx], scale=sc[idx])
