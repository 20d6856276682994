"""Per-role stall summary of a tensor-core kernel from an ncu source-page CSV.
usage: ncu -i rep --page source --csv --print-source sass > x.csv ; python scripts/ncu_stalls.py x.csv [section ...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
secs = []
cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": [], "hdr": None}
        secs.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) >= len(cur["hdr"]) - 2:
        cur["rows"].append(r)
want = [int(a) for a in sys.argv[2:]] or range(len(secs))
for ki in want:
    k = secs[ki]
    h = {n: i for i, n in enumerate(k["hdr"])}
    stalls = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    R = k["rows"]
    S = lambda r: int(r[h["# Samples"]])
    tot = sum(S(r) for r in R)
    src = lambda i: R[i][h["Source"]]
    ld = [i for i in range(len(R)) if "LDTM" in src(i)]
    st = [i for i in range(len(R)) if "UTMASTG" in src(i)]
    mma = [i for i in range(len(R)) if "UTCHMMA" in src(i) or "UTCQMMA" in src(i)]
    f2 = [i for i in range(len(R)) if "F2FP" in src(i) or "LOP3" in src(i)]
    print("== section %d: %d instrs, %d samples; LDTM@%s UTMASTG@%s MMA@%s" % (ki, len(R), tot, ld[:1], st[-1:], mma[:1]))
    if not (ld and st and mma):
        continue
    e0, e1 = ld[0] - 150, st[-1] + 120
    regions = {"setup+producers": (0, 700), "transform": (700, e0), "epilogue": (e0, e1), "mma+tail": (e1, len(R))}
    for nm, (a, b) in regions.items():
        s = sum(S(r) for r in R[a:b])
        ex = sum(int(r[h["Instructions Executed"]]) for r in R[a:b])
        stv = sorted(((sum(int(r[h[n]]) for r in R[a:b]), n[6:]) for n in stalls), reverse=True)[:6]
        print("  %-16s samples %6d (%4.1f%%) warp-instr %10d  %s" % (nm, s, 100.0 * s / max(tot, 1), ex, stv))
    top = sorted(range(len(R)), key=lambda i: -S(R[i]))[:18]
    for i in sorted(top):
        r = R[i]
        tp = sorted(((int(r[h[n]]), n[6:]) for n in stalls), reverse=True)[:2]
        print("    %5d %5.1f%% exec=%9s %-58s %s" % (i, 100.0 * S(r) / max(tot, 1), r[h["Instructions Executed"]], src(i).strip()[:58], tp))
