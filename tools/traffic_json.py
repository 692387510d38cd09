"""ncu --csv log with dram__bytes_{read,write}.sum of one launch -> the JSON bench.py reads (profiles/forward_traffic.json)."""
import csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Metric Name" in r)
mi, vi, ui = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out = {"tokens_per_launch": int(sys.argv[2])}
for r in rows:
    if r is hdr or len(r) <= vi:
        continue
    if r[mi] == "dram__bytes_read.sum":
        out["dram_bytes_read"] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
    if r[mi] == "dram__bytes_write.sum":
        out["dram_bytes_write"] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
print(json.dumps(out))
