"""Post-process an ncu launch list of ONE forward step (bench.py under GRAFP_NCU_RANGE=1,
`ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`):
prints per-kernel-class totals and writes
  <out_prefix>_launches_summary.json   per launch: kernel, grid, us, dram MB
  profiles/ncu_traffic.json            DRAM bytes per segment per kernel class (bench.py's roofline.traffic)
usage: python scripts/ncu_launches.py launches.csv out_prefix segments_per_step [--write-traffic]"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
launch = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    lid = int(r[ix["ID"]])
    e = launch.setdefault(lid, {"kernel": re.sub(r"\(.*", "", r[ix["Kernel Name"]]), "grid": r[ix["Grid Size"]],
                                "block": r[ix["Block Size"]]})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    name = r[ix["Metric Name"]]
    if name.startswith("gpu__time_duration"):
        e["us"] = v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)
    else:
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        e[name.split(".")[0]] = v * scale


def klass(k):
    if "gemm" in k or "ffn_fused" in k:      # the fused FFN is two GEMMs in one launch
        return "gemm"
    if "knn" in k:
        return "knn"
    if "aggregate" in k:
        return "aggregate"
    return "other"


tot = collections.defaultdict(lambda: {"n": 0, "us": 0.0, "bytes": 0.0})
for e in launch.values():
    c = tot[klass(e["kernel"])]
    c["n"] += 1
    c["us"] += e.get("us", 0.0)
    c["bytes"] += e.get("dram__bytes_read", 0.0) + e.get("dram__bytes_write", 0.0)
step_us = sum(c["us"] for c in tot.values())
print("%d launches, %.1f us of kernels (serialised, cold L2, under ncu)" % (len(launch), step_us))
for k, c in sorted(tot.items(), key=lambda kv: -kv[1]["us"]):
    print("  %-10s n=%3d  %9.1f us  %5.1f%% of step  dram %.3f GB" % (k, c["n"], c["us"], 100 * c["us"] / step_us, c["bytes"] / 1e9))
prefix, segs = sys.argv[2], int(sys.argv[3])
json.dump({"segments": segs, "launches": list(launch.values()),
           "class_totals": {k: dict(v, share=v["us"] / step_us) for k, v in tot.items()}},
          open(prefix + "_launches_summary.json", "w"), indent=0)
if "--write-traffic" in sys.argv:
    json.dump({"bytes_per_segment": {k: c["bytes"] / segs for k, c in tot.items() if k != "other"},
               "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, summed over the class's launches of one "
                       "eager forward at %d segments (ncu, %s)" % (segs, os.path.basename(sys.argv[1]))},
              open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
