"""Stall samples / executed instructions of an ncu source-page CSV aggregated by code region (diagnostics).
usage: ncu_regions.py src.csv regions.txt   with lines  name file lo hi"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; cur = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] in ("File Name", "File Path"): cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit(): continue
    try: smp = int(r[hdr.index("# Samples")] or 0); ins = int(r[hdr.index("Instructions Executed")] or 0)
    except ValueError: continue
    agg[(cur, int(r[0]))] = (smp, ins)
tot_s = sum(v[0] for v in agg.values()); tot_i = sum(v[1] for v in agg.values())
print("total samples %d instructions %d" % (tot_s, tot_i))
used = set()
for line in open(sys.argv[2]):
    if not line.strip() or line.startswith("#"): continue
    name, f, lo, hi = line.split(); lo, hi = int(lo), int(hi)
    keys = [k for k in agg if k[0] == f and lo <= k[1] <= hi]
    used.update(keys)
    s = sum(agg[k][0] for k in keys); i = sum(agg[k][1] for k in keys)
    print("%-28s samples %5d (%4.1f%%)  inst %9d (%4.1f%%)" % (name, s, 100 * s / tot_s, i, 100 * i / tot_i))
rest = {}
for k, v in agg.items():
    if k not in used: rest[k[0]] = (rest.get(k[0], (0, 0))[0] + v[0], rest.get(k[0], (0, 0))[1] + v[1])
print("unassigned:", {k: v for k, v in rest.items() if v[0] or v[1]})
