"""Compare per-kernel ms of two bench JSON lines: python scripts/cmp_bench.py a.json b.json"""
import json, sys
a = json.load(open(sys.argv[1])); b = json.load(open(sys.argv[2]))
print("value %.0f -> %.0f   ms/step %.2f -> %.2f   clocks %s -> %s" % (a["value"], b["value"], a["ms_per_step"], b["ms_per_step"], a["clocks"]["sm_mhz"], b["clocks"]["sm_mhz"]))
ka, kb = a["kernel_ms"], b["kernel_ms"]
for k in sorted(set(ka) | set(kb), key=lambda k: -(ka.get(k, 0))):
    x, y = ka.get(k, 0), kb.get(k, 0)
    print("%-48s %7.3f -> %7.3f  %+6.1f%%" % (k, x, y, 100 * (y - x) / x if x else 0))
print("sum %.2f -> %.2f" % (sum(ka.values()), sum(kb.values())))
