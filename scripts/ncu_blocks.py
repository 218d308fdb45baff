"""Group the per-line ncu numbers (scripts/ncu_lines.py input) by the kernel's blocks.
usage: ncu_blocks.py <src.csv> <dis.txt> <func substr> <source file>"""
import csv, re, sys, collections
src_csv, dis, func, srcfile = sys.argv[1:5]
rows = list(csv.reader(open(src_csv))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
lines = []; cur = None; infunc = False
for l in open(dis):
    if l.startswith('.text.'):
        infunc = func in l; continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): lines.append(cur)
# section starts in the source
marks = []
for n, l in enumerate(open(srcfile), 1):
    if '// ----------------' in l or re.match(r'^(__device__|template|__global__|static|inline)', l):
        marks.append((n, l.strip()[:80]))
def section(ln):
    if ln is None or not ln[0].startswith(srcfile.split('/')[-1]): return 'other:' + str(ln[0] if ln else None)
    name = 'top'
    for n, t in marks:
        if n <= ln[1]: name = f"{n}: {t}"
        else: break
    return name
agg = collections.defaultdict(lambda: [0, 0, 0]); tot = [0, 0, 0]
for r, ln in zip(data, lines):
    wi = int(r[ix['Instructions Executed']]); ti = int(r[ix['Thread Instructions Executed']]); sm = int(r[ix['# Samples']])
    a = agg[section(ln)]; a[0] += wi; a[1] += ti; a[2] += sm; tot[0] += wi; tot[1] += ti; tot[2] += sm
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:90s} warp {100*a[0]/tot[0]:5.1f}%  thr/inst {a[1]/max(a[0],1):5.1f}  samples {100*a[2]/tot[2]:5.1f}%")
