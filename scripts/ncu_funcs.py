"""Aggregate an ncu SASS source page by source function ranges (development aid).
usage: ncu_funcs.py <src.csv> <lib.so> <kernel-substring>"""
import csv, re, subprocess, sys, os, tempfile, collections, bisect
srccsv, lib, kern = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {os.path.abspath(lib)} >/dev/null 2>&1", shell=True)
cub = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(f"nvdisasm -g -c {tmp}/{cub}", shell=True, capture_output=True, text=True).stdout.splitlines()
lines, cur, infunc = [], None, False
for ln in dis:
    if ln.startswith('.text.'):
        infunc = kern in ln; continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln): lines.append(cur)
# function start lines per file from the sources
root = os.path.join(os.path.dirname(os.path.abspath(lib)), 'csrc')
funcs = {}
for f in os.listdir(root):
    starts = []
    for i, l in enumerate(open(os.path.join(root, f)), 1):
        m = re.match(r'^(?:template.*\n)?(?:BMPC_DEV|BMPC_HD|BMPC_NOINLINE|__global__|static|inline)\b.*?(\w+)\s*\(', l)
        if m and not l.startswith(' '): starts.append((i, m.group(1)))
    funcs[f] = starts
def fn(file, line):
    st = funcs.get(file)
    if not st: return file
    i = bisect.bisect_right([s[0] for s in st], line) - 1
    return f"{file}:{st[i][1]}" if i >= 0 else file
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]; data = rows[2:]
iS = hdr.index('# Samples'); iB = hdr.index('stall_barrier'); iI = hdr.index('Instructions Executed')
ST = ['stall_no_inst', 'stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_branch_resolving', 'stall_selected']
iST = [hdr.index(x) for x in ST]
assert len(data) == len(lines), (len(data), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0] + [0] * len(ST)); tot = [0, 0, 0]
for r, l in zip(data, lines):
    k = fn(*l) if l else 'none'
    a = agg[k]; s, b, i = int(r[iS]), int(r[iB]), int(r[iI])
    a[0] += s; a[1] += b; a[2] += i; tot[0] += s; tot[1] += b; tot[2] += i
    for q, ix in enumerate(iST): a[3 + q] += int(r[ix])
print('total samples %d barrier %d inst %d' % tuple(tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:40s} smp {100*v[0]/tot[0]:5.1f}% inst {100*v[2]/tot[2]:5.1f}% | bar {100*v[1]/max(1,v[0]):3.0f} " + " ".join(f"{n[6:10]} {100*v[3+q]/max(1,v[0]):3.0f}" for q, n in enumerate(ST)))
