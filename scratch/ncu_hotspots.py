"""Per-source-line hot spots of an `ncu --set full --import-source on` capture (diagnostics).
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src.csv ; python scratch/ncu_hotspots.py /tmp/src.csv [top]
Prints, per CUDA source line, warp stall samples (with the top reasons) and executed warp instructions."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rows = list(csv.reader(open(path)))
fname, hdr = None, None
lines = []          # (file, line, source, samples, inst, reasons dict)
for r in rows:
    if not r:
        continue
    if r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
        continue
    d = dict(zip(hdr, r))
    # hdr has two columns named Source: first = cuda source
    try:
        samples = int(r[hdr.index("# Samples")] or 0)
        inst = int(r[hdr.index("Instructions Executed")] or 0)
    except ValueError:
        continue
    reasons = {}
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            try:
                v = int(d[k] or 0)
            except ValueError:
                v = 0
            if v:
                reasons[k] = v
    lines.append((fname, int(r[0]), r[1].strip()[:100], samples, inst, reasons))
tot_s = sum(x[3] for x in lines)
tot_i = sum(x[4] for x in lines)
by_reason = defaultdict(int)
for x in lines:
    for k, v in x[5].items():
        by_reason[k] += v
print("# total samples %d, executed warp instructions %d" % (tot_s, tot_i))
print("# by reason:", ", ".join("%s %d" % kv for kv in sorted(by_reason.items(), key=lambda kv: -kv[1])[:10]))
print("# ---- by stall samples: file line samples inst | source | top reasons")
for x in sorted(lines, key=lambda x: -x[3])[:top]:
    rs = sorted(x[5].items(), key=lambda kv: -kv[1])[:2]
    print("%s %d %d %d | %s | %s" % (x[0], x[1], x[3], x[4], x[2], rs))
print("# ---- by executed warp instructions")
for x in sorted(lines, key=lambda x: -x[4])[:top]:
    print("%s %d inst %d (%.1f%%) samples %d | %s" % (x[0], x[1], x[4], 100.0 * x[4] / max(tot_i, 1), x[3], x[2]))
