#!/usr/bin/env python3
"""Static per-source-line roll-up of a kernel's SASS (no GPU needed): instructions per source line of OBJECT.o, split by the
pipe that executes them.  usage: sass_lines.py OBJECT.o KERNEL_SUBSTRING [start:end:name ...] [--dump start:end]"""
import collections, os, re, subprocess, sys, tempfile

ALU = ("PRMT", "LOP3", "SHF", "IADD3", "IADD", "ISETP", "FSETP", "FSET", "FMNMX", "FMNMX3", "SEL", "FSEL", "MOV", "LEA", "BMSK", "SGXT", "VIMNMX", "VIMNMX3", "IABS",
       "FCHK", "PLOP3", "P2R", "R2P", "I2FP", "F2FP", "CS2R", "IMNMX", "LOP", "SHL", "SHR", "VIADD", "HMNMX2", "HSETP2", "DSETP", "ISET")
FMA = ("FFMA", "FMUL", "FADD", "IMAD", "FFMA2", "FMUL2", "FADD2", "HFMA2", "HADD2", "HMUL2", "FHFMA", "FHADD")
XU = ("MUFU", "POPC", "FLO", "BREV", "I2F", "F2I", "F2F", "FRND", "I2I")
LSU = ("LDG", "STG", "LDL", "STL", "LDS", "STS", "ATOM", "ATOMG", "ATOMS", "RED", "LDC", "LDCU", "SHFL", "MATCH", "QSPC", "CCTL", "MEMBAR", "ERRBAR")

def cls(op):
    base = op.split(".")[0]
    if base in ALU: return "alu"
    if base in FMA: return "fma"
    if base in XU: return "xu"
    if base in LSU: return "lsu"
    return "ctl"

obj, ksub = sys.argv[1:3]
buckets, dump = [], None
args = sys.argv[3:]
i = 0
while i < len(args):
    if args[i] == "--dump":
        s, e = args[i + 1].split(":"); dump = (int(s), int(e)); i += 2; continue
    s, e, n = args[i].split(":"); buckets.append((int(s), int(e), n)); i += 1
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
per = collections.defaultdict(collections.Counter)
cur, inside, total = None, False, 0
for ln in dis.splitlines():
    if ln.startswith("\t.section\t.text."):
        inside = ksub in ln; continue
    if not inside: continue
    m = re.search(r'//## File ".*?", line (\d+)', ln)
    if m:
        if "inlined at" not in ln: cur = int(m.group(1))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        op = m.group(2)
        per[cur][cls(op)] += 1; per[cur]["all"] += 1; total += 1
        if dump and cur is not None and dump[0] <= cur <= dump[1]:
            print(f"{cur:5d} {ln.strip()[:110]}")
print(f"kernel {ksub}: {total} instructions = {total * 16 / 1024:.1f} KB")
if buckets:
    agg = collections.OrderedDict((n, collections.Counter()) for _, _, n in buckets); agg["other"] = collections.Counter()
    for l, c in per.items():
        name = "other"
        for s, e, n in buckets:
            if l is not None and s <= l <= e: name = n; break
        agg[name].update(c)
    for n, c in agg.items():
        print(f"{n:24s} all {c['all']:5d}  alu {c['alu']:5d}  fma {c['fma']:5d}  xu {c['xu']:4d}  lsu {c['lsu']:4d}  ctl {c['ctl']:4d}")
else:
    for l, c in sorted(per.items(), key=lambda kv: -kv[1]["all"])[:50]:
        print(f"line {l}: all {c['all']:4d} alu {c['alu']:4d} fma {c['fma']:4d} xu {c['xu']:3d} lsu {c['lsu']:3d} ctl {c['ctl']:3d}")
