"""Join an ncu SASS source page (csv) with nvdisasm -g line info: stall samples per source line.
usage: ncu_lines.py <src.csv> <lib.so> <kernel-substring> [top]"""
import csv, re, subprocess, sys, os, tempfile, collections
srccsv, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {os.path.abspath(lib)} >/dev/null 2>&1", shell=True)
cub = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(f"nvdisasm -g -c {tmp}/{cub}", shell=True, capture_output=True, text=True).stdout.splitlines()
# instruction order -> (file, line, inlined chain)
lines, cur, infunc = [], None, False
for ln in dis:
    if ln.startswith('.text.'):
        infunc = kern in ln
        continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3)); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
        lines.append(cur)
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]; data = rows[2:]
iS = hdr.index('# Samples'); iB = hdr.index('stall_barrier'); iI = hdr.index('Instructions Executed')
assert len(data) == len(lines), (len(data), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = 0
for r, l in zip(data, lines):
    s = int(r[iS]); b = int(r[iB]); tot += s
    a = agg[l[:2] if l else None]; a[0] += s; a[1] += b; a[2] += int(r[iI])
print('total samples', tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(k):40s} samples {v[0]:8d} ({100*v[0]/tot:5.1f}%) barrier {v[1]:8d} inst {v[2]}")
