#!/usr/bin/env python
"""Per-kernel DRAM traffic and time from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv` launch list; optionally records the per-launch averages in profiles/traffic.json.

    python scripts/summarize_traffic.py launches.csv [key [slices]]      # key e.g. config4_100k_x_2k/n1/lumped;
                                                                         # slices = PMM_SLICES of the profiled run (default 2)
"""
import collections, csv, json, os, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
iID, iN, iM, iV, iU = hdr.index('ID'), hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}
launch = collections.OrderedDict()
for r in data:
    if len(r) <= iV: continue
    name = r[iN].split('(pmm::')[0].split('(const')[0].strip().replace('void ', '').replace('pmm::', '')
    d = launch.setdefault(r[iID], {"name": name})
    d[r[iM]] = float(r[iV].replace(',', '')) * SCALE.get(r[iU], 1.0)
agg = collections.OrderedDict()
for d in launch.values():
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0.0); a[2] += d.get("dram__bytes_read.sum", 0.0); a[3] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | share | DRAM read GB | DRAM write GB | GB/s |\n|---|---|---|---|---|---|---|")
for k, (c, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f%% | %.3f | %.3f | %.0f |" % (k[:70], c, t, 100 * t / tot, rd / 1e9, wr / 1e9, (rd + wr) / max(t, 1e-9) / 1e3))
print("\ntotal: %d launches, %.1f us (cold-cache, serialised under ncu: compare shares, not absolutes)" % (len(launch), tot))
if len(sys.argv) > 2:
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    tj = json.load(open(path)) if os.path.exists(path) else {}
    tj[sys.argv[2]] = {re.match(r'\w+', k).group(0): (rd + wr) / c for k, (c, t, rd, wr) in agg.items() if k.startswith("pmm_l")}
    tj[sys.argv[2]]["_slices"] = int(sys.argv[3]) if len(sys.argv) > 3 else 2      # a sliced launch covers 1/slices of the sites
    tj[sys.argv[2]]["_note"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch (average over the launches of %s)" % os.path.basename(sys.argv[1])
    json.dump(tj, open(path, "w"), indent=1, sort_keys=True)
