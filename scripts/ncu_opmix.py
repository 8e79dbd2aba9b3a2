#!/usr/bin/env python
"""dynamic opcode mix of a kernel from an .ncu-rep (warp instructions executed per opcode)"""
import csv, subprocess, sys, re, collections
rep = sys.argv[1]; steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
data = [r for r in rows if r and r[0].startswith("0x")]
ops = collections.Counter(); tot = 0
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[1])
    ex = int(r[5]); tot += ex
    ops[m.group(2)] += ex
print("total warp-instr", tot, ("per warp-step %.1f" % (tot / steps)) if steps else "")
for k, v in ops.most_common(30):
    print("  %-10s %12d %5.1f%%" % (k, v, 100 * v / tot), ("%.1f/step" % (v / steps)) if steps else "")
