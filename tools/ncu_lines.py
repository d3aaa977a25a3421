#!/usr/bin/env python3
"""Per-source-line roll-up of an ncu capture: joins the SASS rows of `ncu --page source --csv` with the line table of the
matching cubin (nvdisasm -g).  usage: ncu_lines.py REPORT.ncu-rep OBJECT.o KERNEL_SUBSTRING [start:end:name ...]"""
import csv, io, re, subprocess, sys, os, tempfile, collections

rep, obj, ksub = sys.argv[1:4]
buckets = []
for b in sys.argv[4:]:
    s, e, n = b.split(":")
    buckets.append((int(s), int(e), n))

tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# line table of the kernel
lines, cur, inside = {}, None, False
for ln in dis.splitlines():
    if ln.startswith("\t.section\t.text."):
        inside = ksub in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File ".*?", line (\d+)', ln)
    if m:
        cur = int(m.group(1))   # innermost frame comes first; "inlined at" lines follow and are ignored below
        continue
    if "inlined at" in ln:
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        lines[int(m.group(1), 16)] = cur

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr_i]
ci = {k: h.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
base = None
per_line = collections.defaultdict(lambda: [0, 0, 0])
for r in rows[hdr_i + 1:]:
    if len(r) < len(h):
        continue
    a = int(r[ci["Address"]], 16)
    if base is None:
        base = a
    ln = lines.get(a - base)
    v = per_line[ln]
    v[0] += int(r[ci["Instructions Executed"]] or 0)
    v[1] += int(r[ci["Thread Instructions Executed"]] or 0)
    v[2] += int(r[ci["# Samples"]] or 0)
tot = [sum(v[i] for v in per_line.values()) for i in range(3)]
print(f"total: warp-instr {tot[0]/1e9:.2f} G, thread-instr {tot[1]/1e9:.1f} G, lanes {tot[1]/max(tot[0],1):.2f}, samples {tot[2]}")
if buckets:
    agg = collections.OrderedDict((n, [0, 0, 0]) for _, _, n in buckets)
    agg["other"] = [0, 0, 0]
    for ln, v in per_line.items():
        name = "other"
        for s, e, n in buckets:
            if ln is not None and s <= ln <= e:
                name = n
                break
        for i in range(3):
            agg[name][i] += v[i]
    for n, v in agg.items():
        print(f"{n:28s} warp-instr {100*v[0]/tot[0]:5.1f} %  thread-instr {100*v[1]/tot[1]:5.1f} %  lanes {v[1]/max(v[0],1):5.2f}  samples {100*v[2]/max(tot[2],1):5.1f} %")
else:
    for ln, v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:60]:
        print(f"line {ln}: warp-instr {100*v[0]/tot[0]:5.2f} %  lanes {v[1]/max(v[0],1):5.2f}  samples {100*v[2]/max(tot[2],1):5.2f} %")
