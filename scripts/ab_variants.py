"""A/B of k_solve builds on the bench workload (development aid): for every library given, device throughput on B
instances and the outputs, compared with those of the first library.
usage: python scripts/ab_variants.py B lib1.so lib2.so ...   (each library runs in its own process via BMPC_LIB)"""
import os, sys, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 2 and sys.argv[1] != "--child":
    B = sys.argv[1]
    outs = []
    for lib in sys.argv[2:]:
        out = os.path.join(ROOT, "gpurun_out", "ab_" + os.path.basename(lib).replace(".so", "") + ".npz")
        env = dict(os.environ, BMPC_LIB=os.path.abspath(lib))
        subprocess.run([sys.executable, __file__, "--child", B, out], env=env, check=False)
        outs.append(out)
    import numpy as np
    ref = np.load(outs[0])
    for o in outs[1:]:
        if not os.path.exists(o):
            print(json.dumps({"lib": o, "error": "no output"})); continue
        d = np.load(o)
        same = (d["x"] == ref["x"]).all(axis=1)
        ok = (d["status"] == 0) & (ref["status"] == 0)
        rel = np.abs(d["x"] - ref["x"]).max(axis=1) / np.maximum(1.0, np.abs(ref["x"]).max(axis=1))
        print(json.dumps({"lib": os.path.basename(o), "bitwise_equal_instances": int(same.sum()), "of": int(same.size),
                          "status_equal": int((d["status"] == ref["status"]).sum()), "iters_equal": int((d["iters"] == ref["iters"]).sum()),
                          "max_rel_dx_both_ok": float(rel[ok].max()), "n_rel_gt_1e-6": int((rel[ok] > 1e-6).sum()),
                          "f_rel_max": float((np.abs(d["f"] - ref["f"]) / np.maximum(1e-12, np.abs(ref["f"])))[ok & (rel < 1e-6)].max())}), flush=True)
    sys.exit(0)
sys.path.insert(0, ROOT)
import numpy as np, torch
from boundmpc_b200.ocp import default_solver
from boundmpc_b200 import batches
B, outp = int(sys.argv[2]), sys.argv[3]
s = default_solver(solver_opts={'b200': {'tol': float(os.environ.get('AB_TOL', '1e-5'))}})
x0, p = batches.make_batch(default_solver(), ("exp1", "exp2"), int(os.environ.get("AB_FIRST", "0")), B, bound_scale=True)
xd, pd = torch.from_numpy(x0).cuda(), torch.from_numpy(p).cuda()
out = s.solve_batch(xd, pd); torch.cuda.synchronize()
best = 1e9
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.solve_batch(xd, pd, out); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"lib": os.path.basename(os.environ.get("BMPC_LIB", "default")), "B": B, "ms": best, "solves_per_s": B / best * 1e3,
                  "ok": int((out["status"] == 0).sum()), "iters_mean": float(out["iters"].double().mean()), "iters_max": int(out["iters"].max())}), flush=True)
np.savez(outp, x=out["x"].cpu().numpy(), f=out["f"].cpu().numpy(), iters=out["iters"].cpu().numpy(), status=out["status"].cpu().numpy())
