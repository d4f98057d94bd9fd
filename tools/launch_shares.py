#!/usr/bin/env python
"""Shares of device time per kernel from an ncu launch list (--metrics gpu__time_duration.sum --csv): launch_shares.py <csv> [title]"""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]; ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
ui = h.index("Metric Unit")
t = collections.OrderedDict()
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum": continue
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}[r[ui]]
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    a = t.setdefault(name, [0.0, 0]); a[0] += v; a[1] += 1
tot = sum(a[0] for a in t.values()); n = sum(a[1] for a in t.values())
if len(sys.argv) > 2: print("# " + sys.argv[2])
print(f"# {n} launches, {tot:.2f} ms of device time under ncu (cold-cache, serialised: compare shares)")
for k, a in sorted(t.items(), key=lambda kv: -kv[1][0]):
    print(f"{100 * a[0] / tot:6.2f} %  {a[0]:10.3f} ms  {a[1]:5d} launches  {k}")
