"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i + 1
        break
ki, mi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0
for r in rows[start:]:
    if len(r) <= mi:
        continue
    name = re.sub(r'\(.*', '', r[ki])
    v = float(r[mi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:10.1f} us {n:5d} {100 * t / tot:5.1f}%  {k[:110]}")
print(f"{tot:10.1f} us total")
