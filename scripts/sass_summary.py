"""SASS instruction mix of every k_solve variant (spill regressions and the FP64 tensor path at a glance):
python scripts/sass_summary.py > profiles/sass_summary.txt"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "boundmpc_b200", "libboundmpc_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, cnt = None, collections.defaultdict(collections.Counter)
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1).split(".")[0]
        cnt[cur][op] += 1
        cnt[cur]["_total"] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: static instruction counts per kernel (sm_100a)")
print(f"{'kernel':34s} {'total':>7s} {'KB':>6s} {'DMMA':>6s} {'DFMA':>6s} {'DMUL':>6s} {'DADD':>6s} {'MUFU':>5s} {'LDGSTS':>6s} {'LDS':>6s} {'STS':>6s} {'LDG':>5s} {'STG':>5s} {'LDL':>5s} {'STL':>5s} {'BAR':>4s} {'SHFL':>5s}")
for k in sorted(cnt):
    c = cnt[k]
    print(f"{k[:34]:34s} {c['_total']:7d} {c['_total'] * 16 / 1024:6.1f} {c['DMMA']:6d} {c['DFMA']:6d} {c['DMUL']:6d} {c['DADD']:6d} {c['MUFU']:5d} {c['LDGSTS']:6d} "
          f"{c['LDS']:6d} {c['STS']:6d} {c['LDG']:5d} {c['STG']:5d} {c['LDL']:5d} {c['STL']:5d} {c['BAR']:4d} {c['SHFL']:5d}")
