"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown summary of
ONE V-cycle (the launches between two level-0 down-leg kernels).
usage: python profiles/summarize_launches.py launches.csv 'first-kernel-substring' > out.md"""
import csv
import sys
from collections import defaultdict

path, first = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
iK, iV, iG = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
L = [(r[iK], float(r[iV].replace(",", "")) / 1000.0, r[iG]) for r in rows[1:]]
starts = [i for i, l in enumerate(L) if first in l[0]]
period = starts[1] - starts[0] if len(starts) > 1 and starts[1] - starts[0] > 3 else len(L) - starts[0]
cyc = L[starts[0]:starts[0] + period]
tot = sum(x[1] for x in cyc)
agg, cnt = defaultdict(float), defaultdict(int)
for k, v, g in cyc:
    name = k.split("(")[0].replace("void ", "").replace("amgb::", "")
    agg[name] += v
    cnt[name] += 1
print("one V-cycle = %d launches, %.1f us summed (ncu serialises the launches and runs them cold: compare shares)\n" % (len(cyc), tot))
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k in sorted(agg, key=lambda k: -agg[k]):
    print("| %s | %d | %.1f | %.1f%% |" % (k, cnt[k], agg[k], 100 * agg[k] / tot))
print("\n| # | kernel | us | grid |\n|---|---|---|---|")
for i, (k, v, g) in enumerate(cyc):
    print("| %d | %s | %.1f | %s |" % (i, k.split("(")[0].replace("void ", "").replace("amgb::", ""), v, g))
