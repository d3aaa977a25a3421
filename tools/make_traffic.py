#!/usr/bin/env python3
"""profiles/traffic.json from ncu --set full reports: per kernel of each report ONE launch's DRAM bytes, L2 bytes, thread / warp instructions,
duration, issue-slot and DRAM / L2 utilisation.  usage: make_traffic.py WORKLOAD:nN:REPORT.ncu-rep [...]   (merges into the existing file)"""
import csv, io, json, os, re, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
path = os.path.join(ROOT, "profiles", "traffic.json")
table = json.load(open(path)) if os.path.exists(path) else {}
table = {k: v for k, v in table.items() if isinstance(v, dict) or k.startswith("_")}
UNIT = {"sector": 1, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "inst": 1, "": 1, "ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6, "%": 1, "thread": 1}
for spec in sys.argv[1:]:
    wl, nn, rep = spec.split(":", 2)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        def val(k):
            return float(d[k].replace(",", "")) * UNIT.get(u[k], 1) if d.get(k) not in (None, "", "n/a") else None
        name = re.sub(r"<.*", "", d["Kernel Name"].split("(")[0].split("::")[-1]).strip()
        rec = {"dram_bytes": (val("dram__bytes_read.sum") or 0) + (val("dram__bytes_write.sum") or 0), "dram_read": val("dram__bytes_read.sum"),
               "dram_write": val("dram__bytes_write.sum"), "lts_bytes": (val("lts__t_sectors.sum") or 0) * 32.0,
               "thread_inst": (val("smsp__inst_executed.sum") or 0) * (val("smsp__thread_inst_executed_per_inst_executed.ratio") or 0),
               "warp_inst": val("smsp__inst_executed.sum"), "duration_ms_under_ncu": val("gpu__time_duration.sum"),
               "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "dram_pct_of_peak": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
               "l2_pct_of_peak": val("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
               "l1_pct_of_peak": val("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), "report": os.path.basename(rep)}
        table[f"{name}:{wl}:{nn}"] = rec   # the last launch of a kernel in the report wins
table["_source"] = ("ncu --set full --clock-control none under gpurun (reports stay in gpurun_out/, one launch per kernel); written by tools/make_traffic.py. "
                    "dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum, lts_bytes = 32 B x lts__t_sectors.sum, thread_inst = smsp__inst_executed.sum x smsp__thread_inst_executed_per_inst_executed.ratio")
json.dump(table, open(path, "w"), indent=1)
print("\n".join(f"{k}: {v}" for k, v in table.items() if not k.startswith("_")))
