#!/usr/bin/env python3
"""SASS listing of a kernel from an ncu capture, in address order, with the source line, executed warp instructions and active lanes per
instruction.  usage: ncu_sass.py REPORT.ncu-rep OBJECT.o KERNEL_SUBSTRING [firstLine:lastLine ...]   (only instructions of those lines)"""
import csv, io, re, subprocess, sys, os, tempfile
rep, obj, ksub = sys.argv[1:4]
ranges = [tuple(int(v) for v in a.split(":")) for a in sys.argv[4:]]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines, cur, inside = {}, None, False
for ln in dis.splitlines():
    if ln.startswith("\t.section\t.text."):
        inside = ksub in ln; continue
    if not inside: continue
    m = re.search(r'//## File ".*?", line (\d+)', ln)
    if m:
        if "inlined at" not in ln: cur = int(m.group(1))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m: lines[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ci = {k: h.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
base = None
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    a = int(r[ci["Address"]], 16)
    if base is None: base = a
    ln = lines.get(a - base)
    if ranges and not any(ln is not None and s <= ln <= e for s, e in ranges): continue
    n = int(r[ci["Instructions Executed"]] or 0); t = int(r[ci["Thread Instructions Executed"]] or 0)
    print(f"{a - base:6x} L{ln if ln is not None else 0:<5d} {n / 1e6:8.2f}M {t / max(n, 1):5.1f} {int(r[ci['# Samples']] or 0):6d}  {r[ci['Source']].strip()[:90]}")
