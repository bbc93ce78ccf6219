"""Small batch through every path of k_solve for compute-sanitizer (development aid): golden sequences (warm starts),
literal-spec fixtures (cold-start repair, infeasible crawls, second-order corrections), config-4 fixtures (N = 20)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from boundmpc_b200.ocp import default_solver
from tests.util import load
S1, S2, SC, C4 = load("seq_exp1.npz"), load("seq_exp2.npz"), load("spec_cold.npz"), load("cfg4_tight.npz")
x0 = np.concatenate([S1["x0"][:6], S2["x0"][:4], SC["x0"]]); p = np.concatenate([S1["p"][:6], S2["p"][:4], SC["p"]])
s = default_solver()
r = s.solve_batch(x0, p)
print("N=10", r["iters"].tolist(), r["status"].tolist())
s20 = default_solver(N=20)
r = s20.solve_batch(C4["x0"][[0, 3]], C4["p"][[0, 3]])
print("N=20", r["iters"].tolist(), r["status"].tolist())
