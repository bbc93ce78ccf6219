"""Copy the outputs of scripts/gpu_r2_final.sh (gpurun_out/r2) into profiles/ with the round prefix and derive the
summaries (ncu metrics, per-function stalls, DRAM traffic).  usage: python scripts/collect_profiles.py [src] [prefix]"""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r2")
pre = sys.argv[2] if len(sys.argv) > 2 else "r2"
P = os.path.join(ROOT, "profiles")
def cp(a, b):
    if os.path.exists(os.path.join(src, a)):
        shutil.copy(os.path.join(src, a), os.path.join(P, f"{pre}_{b}")); print("copied", b)
for a, b in [("bench.json", "bench.json"), ("bench_reference.json", "bench_reference.json"), ("shard_sweep.json", "shard_sweep.json"),
             ("trace_util_s0.txt", "trace_util_shard0.txt"), ("trace_util_s4.txt", "trace_util_shard4.txt"),
             ("phase_cycles_b148.txt", "phase_cycles_b148.txt"), ("phase_cycles_b8192.txt", "phase_cycles_b8192.txt"),
             ("launches_bench.csv", "launches_bench.csv"), ("env_ab_l2.log", "ab_l2_window.txt"), ("pytest_gpu.log", "pytest_gpu.log")]:
    cp(a, b)
for cfg in ("exp1_1024", "exp2_8192", "exp1_N20_tight_8192", "spec_mixed_65536"):
    cp(f"bench_{cfg}.json", f"bench_{cfg}.json")
hdr = "# ncu --set full --clock-control none of ONE k_solve launch (8,192 instances, 1 B200, tol = AB_TOL of scripts/profile_batch.py: 1e-5 from r2b on), final code of the round; numbers under the profiler are not bench values\n"
for raw, out, what in [("prof_raw.csv", "k_solve_metrics.txt", "mixed_65536 shard"), ("prof_n20_raw.csv", "k_solve_N20_metrics.txt", "exp1_N20_tight_8192")]:
    f = os.path.join(src, raw)
    if os.path.exists(f):
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), f], capture_output=True, text=True).stdout
        open(os.path.join(P, f"{pre}_{out}"), "w").write(hdr + f"# workload: {what}\n" + txt); print("wrote", out)
f = os.path.join(src, "prof_source.csv")
if os.path.exists(f):
    lib = os.path.join(ROOT, "boundmpc_b200", "libboundmpc_b200.so")
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_funcs.py"), f, lib, "k_solveILi128ELi3"], capture_output=True, text=True)
    open(os.path.join(P, f"{pre}_k_solve_by_function.txt"), "w").write("# warp-state samples of k_solve<128,3> by source function (ncu source page joined with nvdisasm line info)\n" + txt.stdout + txt.stderr[-500:])
    print("wrote by_function")
tr = {}
for w, name in ((0, "with_l2_window"), (1, "without_l2_window"), (2, "default")):
    f = os.path.join(src, f"dram_nowindow{w}.csv" if w < 2 else "dram.csv")
    if os.path.exists(f):
        rows = [r for r in csv.reader(open(f)) if len(r) > 5]
        h = rows[0]
        iN, iV = h.index("Metric Name"), h.index("Metric Value")
        tr[name] = {r[iN]: float(r[iV].replace(",", "")) for r in rows[1:] if "k_solve" in " ".join(r)}
if tr:
    d = tr.get("default") or tr.get("with_l2_window", {})
    tot = None
    if d:
        rd, wr = d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)
        tot = rd + wr
    json.dump({"k_solve_dram_bytes_per_launch": tot, "unit_note": "ncu units as printed (see per-setting dicts)", "settings": tr,
               "algorithmic_bytes_per_launch": 8192 * 21504}, open(os.path.join(P, f"{pre}_traffic.json"), "w"), indent=1)
    print("wrote traffic", tr)
