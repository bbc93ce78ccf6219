"""Static code size of a kernel by source function (development aid).
usage: code_size.py <lib.so> <kernel-substring>"""
import re, subprocess, sys, os, tempfile, collections, bisect
lib, kern = sys.argv[1:3]
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {os.path.abspath(lib)} >/dev/null 2>&1", shell=True)
cub = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(f"nvdisasm -g -c {tmp}/{cub}", shell=True, capture_output=True, text=True).stdout.splitlines()
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'boundmpc_b200', 'csrc')
funcs = {}
for f in os.listdir(root):
    starts = []
    for i, l in enumerate(open(os.path.join(root, f)), 1):
        m = re.match(r'^(?:BMPC_DEV|BMPC_HD|BMPC_NOINLINE|__global__|static|inline)\b.*?(\w+)\s*\(', l)
        if m and not l.startswith(' '): starts.append((i, m.group(1)))
    funcs[f] = starts
def fn(file, line):
    st = funcs.get(file)
    if not st: return file
    i = bisect.bisect_right([s[0] for s in st], line) - 1
    return f"{file}:{st[i][1]}" if i >= 0 else file
cnt = collections.Counter(); cur = None; infunc = False; sect = None; per_sect = collections.Counter()
for ln in dis:
    if ln.startswith('.text.') or ln.startswith('\t.section\t.text.') or re.match(r'\s*\.section\s+\.text\.', ln):
        infunc = kern in ln; sect = ln.strip(); continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
        cnt[fn(*cur) if cur else 'none'] += 1; per_sect[sect] += 1
tot = sum(cnt.values())
print("total %.1f KB" % (tot * 16 / 1024))
for k, v in per_sect.most_common(12): print("  section %-90s %.1f KB" % (k[:90], v * 16 / 1024))
for k, v in cnt.most_common(30): print(f"{k:45s} {v*16/1024:7.1f} KB")
