"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel.
    python tools/launch_summary.py gpurun_out/launches.csv [--md]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
md = "--md" in sys.argv
if md:
    print("| share | total us | launches | avg us | kernel |\n|---|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if md:
        print(f"| {t / tot * 100:.2f}% | {t:.1f} | {n} | {t / n:.1f} | `{k[:90]}` |")
    else:
        print(f"{t / tot * 100:6.2f}%  {t:10.1f} us  n={n:3d}  avg {t / n:9.1f} us  {k[:90]}")
print(f"\ntotal {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
