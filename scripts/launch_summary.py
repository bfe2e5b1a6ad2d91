"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count, total and
mean device time and share of the listed launches (shares, not absolutes, are comparable with bench.py's
CUDA-event breakdown: ncu serialises launches and runs them cold-cache).

    python scripts/launch_summary.py gpurun_out/launches.csv [last_n_launches] > profiles/rNN_launches.txt
"""
import csv
import re
import sys

path = sys.argv[1]
last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
for r in rd:
    if r[ix['Metric Name']] != 'gpu__time_duration.sum':
        continue
    name = r[ix['Kernel Name']]
    name = re.sub(r'\(.*', '', name)
    name = re.sub(r'^void ', '', name)
    if name.startswith('at::') or 'at::native' in name:
        name = 'torch:' + name.split('<')[0].split('::')[-1]
    val = float(r[ix['Metric Value']].replace(',', ''))
    unit = r[ix['Metric Unit']]
    ns = val * {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1.0)
    rows.append((name[:70], r[ix['Grid Size']], r[ix['Block Size']], ns))
if last:
    rows = rows[-last:]
tot = sum(r[3] for r in rows)
agg = {}
for n, g, b, ns in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += ns
print('launches listed: %d   total device time: %.3f ms' % (len(rows), tot / 1e6))
print('%-70s %6s %10s %10s %7s' % ('kernel', 'n', 'total us', 'mean us', 'share'))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-70s %6d %10.1f %10.2f %6.1f%%' % (n, c, t / 1e3, t / 1e3 / c, 100 * t / tot))
