"""Summarise an `ncu --page source --csv` dump: top stalled SASS instructions with their stall reasons."""
import csv, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for n, r in enumerate(rows[2:]):
    try: s = int(r[ci["# Samples"]])
    except Exception: continue
    top = sorted(((int(r[ci[h]] or 0), h) for h in stalls), reverse=True)[:2]
    data.append((s, n, r[ci["Source"]].strip()[:90], int(r[ci["Instructions Executed"]] or 0), top))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for s, n, src, ex, top in sorted(data, reverse=True)[:topn]:
    print(f"{s:7d} {100*s/tot:5.1f}%  #{n:5d} exec={ex:9d}  {src:90s} {top}")
