"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line (instructions executed and
stall samples). Usage: python profiles/agg_source.py dump.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg, cur, fname = {}, None, ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ie:
        continue
    if r[0] != "":
        try:
            cur = (fname, int(r[0]), r[1])
        except ValueError:
            continue
    if r[2] == "":
        continue
    try:
        n, s = int(r[ie]), int(r[isamp])
    except ValueError:
        continue
    a = agg.setdefault(cur, [0, 0])
    a[0] += n
    a[1] += s
tot = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print("total warp-instructions", tot, "samples", ts)
for (f, ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n:10d} {100 * n / tot:5.1f}%  samples {100 * s / ts:5.1f}%  {f}:{ln:<4d} {src.strip()[:100]}")
