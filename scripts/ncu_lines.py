"""Join ncu's SASS-level source page with nvdisasm line info: per-source-line instruction and sample totals.

usage: ncu_lines.py <ncu --page source --csv output> <nvdisasm -g -c output> <mangled function substring> [top]
"""
import csv, re, sys, collections
src_csv, dis, func = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
# instruction -> line from nvdisasm
lines = []
cur = None; infunc = False
for l in open(dis):
    if l.startswith('.text.'):
        infunc = func in l
        continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        lines.append(cur)
print('sass instrs: ncu', len(data), 'nvdisasm', len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r, ln in zip(data, lines):
    wi = int(r[ix['Instructions Executed']]); ti = int(r[ix['Thread Instructions Executed']]); sm = int(r[ix['# Samples']])
    a = agg[ln]; a[0] += wi; a[1] += ti; a[2] += sm
    tot[0] += wi; tot[1] += ti; tot[2] += sm
print('total warp inst', tot[0], 'thread inst', tot[1], 'avg thr', tot[1] / tot[0], 'samples', tot[2])
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(ln):40s} warp {100*a[0]/tot[0]:5.1f}%  thr/inst {a[1]/max(a[0],1):5.1f}  samples {100*a[2]/tot[2]:5.1f}%")
