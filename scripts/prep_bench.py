"""Device time of k_prepare / k_post on 8,192 tiled instances (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches
s = default_solver()
D = batches.make_builder_batch(s, ("exp1", "exp2"), 0, 256, bound_scale=True)
B = 8192
tile = lambda a: np.ascontiguousarray(np.concatenate([a] * (B // 256)))
dev = torch.device("cuda")
t = {k: torch.from_numpy(tile(D[k])).to(dev) for k in ("path_id", "sector", "state", "prev", "sector_out", "x0")}
tabs = torch.from_numpy(D["tables"]).to(dev)
sec0 = t["sector"].clone()
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
bo = s.prepare_batch(tabs, t["path_id"], t["sector"], t["state"], t["prev"])
def prep():
    t["sector"].copy_(sec0); s.prepare_batch(tabs, t["path_id"], t["sector"], t["state"], t["prev"], bo)
po = s.post_batch(tabs, t["path_id"], t["sector_out"], t["state"], t["x0"])
def post():
    s.post_batch(tabs, t["path_id"], t["sector_out"], t["state"], t["x0"], None, po)
print(f"k_prepare {timeit(prep):.4f} ms   k_post {timeit(post):.4f} ms   (B = {B})")
