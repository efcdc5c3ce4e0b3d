"""Split ncu's SASS-level stall samples of the chunk kernel into phases delimited by marker
instructions (BAR.SYNC / mbarrier waits / UTCBAR).  Usage: ncu_segments.py <source.csv>"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; si = hdr.index('# Samples'); src = hdr.index('Source')
stall_cols = [(n, hdr.index(n)) for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
data = []
for r in rows[2:]:
    if len(r) <= si or not r[si].isdigit(): continue
    data.append((int(r[si]), r[src].strip(), {n: int(r[i] or 0) for n, i in stall_cols}))
tot = sum(d[0] for d in data)
print('total samples', tot, 'instructions', len(data))
prev_i, prev_a, acc = 0, 0, 0
segs = []
for i, (s, t, st) in enumerate(data):
    acc += s
    if re.search(r'BAR\.SYNC|UTCBAR|SYNCS\.PHASECHK', t):
        segs.append((prev_i, i, acc - prev_a, t[:48])); prev_i, prev_a = i + 1, acc
segs.append((prev_i, len(data), acc - prev_a, 'end'))
for a, b, s, t in segs:
    if s / tot < 0.003: continue
    c = collections.Counter()
    for x in data[a:b + 1]:
        for n, v in x[2].items(): c[n] += v
    top = ', '.join(f"{n[6:]} {100*v/max(1,s):.0f}%" for n, v in c.most_common(3))
    print(f"[{a:5d}-{b:5d}] {100*s/tot:5.1f}%  ends at {t:48s} | {top}")
