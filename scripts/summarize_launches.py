#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
iN, iV, iU = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in data:
    if len(r) <= iV: continue
    name = r[iN].split('(pmm::KParams')[0].split('(const')[0].strip()
    v = float(r[iV].replace(',', ''))
    v = v / 1000.0 if r[iU] == 'ns' else (v * 1000.0 if r[iU] == 'ms' else v)
    a = agg.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | share | max us |\n|---|---|---|---|---|")
for k, (c, t, m) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f%% | %.1f |" % (k[:90], c, t, 100 * t / tot, m))
print("\ntotal: %d launches, %.1f us (cold-cache, serialised under ncu: compare shares, not absolutes)" % (len(data), tot))
