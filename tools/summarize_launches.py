"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean
duration and share.  usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/x.md"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    agg.setdefault(r[ki].split("(")[0][-60:], []).append(v)
tot = sum(sum(v) for v in agg.values())
print("| kernel | launches | mean us | share |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"| `{k}` | {len(v)} | {sum(v) / len(v):.2f} | {100 * sum(v) / tot:.1f}% |")
