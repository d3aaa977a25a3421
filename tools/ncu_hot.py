#!/usr/bin/env python3
"""Hottest SASS instructions of an ncu capture by stall samples, with the dominant stall reason, + totals per reason and a listing
of taken-branch style 'no_inst' hot spots.  usage: ncu_hot.py REPORT.ncu-rep [N]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
st = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
ix = {c: h.index(c) for c in st + ["Address", "Source", "# Samples", "Instructions Executed", "Avg. Threads Executed"]}
data = []
tot = {c: 0 for c in st}
for k, r in enumerate(rows[hi + 1:]):
    if len(r) < len(h): continue
    s = int(r[ix["# Samples"]] or 0)
    d = {c: int(r[ix[c]] or 0) for c in st}
    for c in st: tot[c] += d[c]
    data.append((k, s, r[ix["Source"]].strip(), d, int(r[ix["Instructions Executed"]] or 0), r[ix["Avg. Threads Executed"]]))
allS = sum(x[1] for x in data)
print("samples", allS, " by reason:", ", ".join(f"{c[6:]} {100*v/allS:.1f}%" for c, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 200 > allS))
for k, s, src, d, ie, th in sorted(data, key=lambda x: -x[1])[:N]:
    top = sorted(d.items(), key=lambda kv: -kv[1])[:2]
    print(f"#{k:5d} {100*s/allS:5.2f}%  exec {ie/1e6:7.2f}M thr {th:>5s}  {src[:60]:60s} " + " ".join(f"{c[6:]}={v}" for c, v in top if v))
