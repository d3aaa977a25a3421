import sys, os, json
sys.path.insert(0, os.getcwd())
import raygun_b200 as rg, bench
o = bench.measure_also(rg, "c4", 0)
print(os.environ.get("RGB200_LIB", "default"), o["ms_per_step"], o["value"], o["sections_ms"])
