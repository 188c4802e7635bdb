"""Summarise an ncu launch list (gpu__time_duration.sum CSV) of `bench.py --steps 1 --warmup 1`: per-kernel totals of ONE step
(the last complete step, delimited by pack_conv_weight_kernel launches).  Usage: python scripts/launch_summary.py <csv> [-v]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
verbose = len(sys.argv) > 2
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ik, iv, ig = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
data = [(x[ik], float(x[iv]), x[ig]) for x in r]
starts = [i for i, (k, v, g) in enumerate(data) if 'pack_conv_weight' in k]
a, b = starts[-2], starts[-1]
step = data[a:b]
tot = sum(v for k, v, g in step)
print("launches/step", len(step), "sum of kernel durations %.1f us" % (tot / 1e3))
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v, g in step:
    k = re.sub(r'\(.*', '', k).replace('void ', '')
    agg[k][0] += 1
    agg[k][1] += v
for k, (n_, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print("%8.1f us %5.1f%% n=%3d avg %7.1f  %s" % (v / 1e3, 100 * v / tot, n_, v / n_ / 1e3, k[:80]))
if verbose:
    for i, (k, v, g) in enumerate(step):
        print(i, re.sub(r'\(.*', '', k).replace('void ', '').replace('drn::', '')[:44], g, "%.1f" % (v / 1e3))
