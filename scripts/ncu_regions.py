"""Histogram of stall samples / executed instructions per block of SASS instructions.
usage: python scripts/ncu_regions.py source_page.csv [block=150] [section=0]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 150
sec = int(sys.argv[3]) if len(sys.argv) > 3 else 0
secs = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"rows": [], "hdr": None}; secs.append(cur)
    elif cur is not None and r and r[0] == "Address": cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) >= len(cur["hdr"]) - 2: cur["rows"].append(r)
k = secs[sec]; h = {n: i for i, n in enumerate(k["hdr"])}; R = k["rows"]
stalls = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
S = lambda r: int(r[h["# Samples"]]); tot = sum(S(r) for r in R)
src = lambda i: R[i][h["Source"]].strip()
ex = lambda i: int(R[i][h["Instructions Executed"]])
print("instrs", len(R), "samples", tot, "warp-instr executed %.1fM" % (sum(ex(i) for i in range(len(R))) / 1e6))
for a in range(0, len(R), blk):
    b = min(len(R), a + blk)
    s = sum(S(r) for r in R[a:b]); e = sum(ex(i) for i in range(a, b))
    if s < 0.01 * tot and e < 1e6: continue
    stv = sorted(((sum(int(r[h[n]]) for r in R[a:b]), n[6:]) for n in stalls), reverse=True)[:4]
    c = collections.Counter()
    for i in range(a, b):
        op = re.sub(r'^@!?U?P\d+\s+', '', src(i)).split()[0].split('.')[0]; c[op] += ex(i)
    print("%5d samples %5.1f%% exec %6.1fM %s %s" % (a, 100.0 * s / tot, e / 1e6, stv, [(o, round(n / 1e6, 1)) for o, n in c.most_common(5)]))
