"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import csv
import re
import sys

path, out = sys.argv[1], sys.argv[2]
steps = float(sys.argv[3]) if len(sys.argv) > 3 else None
lines = [l for l in open(path) if not l.startswith('==')]
agg = {}
total = 0.
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('<unnamed>::', '').strip()
    try:
        t = float(row['Metric Value'].replace(',', '')) / 1000.
    except ValueError:
        continue
    if t != t:
        continue
    a = agg.setdefault(name, [0., 0])
    a[0] += t
    a[1] += 1
    total += t
with open(out, 'w') as f:
    f.write('| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n')
    for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        f.write('| `%s` | %d | %.0f | %.1f %% | %.1f |\n' % (k[:90], n, t, 100 * t / total, t / n))
    f.write('\ntotal %.0f us over %d launches (ncu: serialised, cold caches; shares are meaningful, absolutes are not)\n' % (total, sum(n for _, n in agg.values())))
print(open(out).read())
