"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (and grid size)."""
import collections, csv, sys
path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr = r; rows = rows[i + 1:]; break
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
by = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[ki].split("(")[0]
    key = (name, r[gi]) if "--grid" in sys.argv else (name,)
    by[key][0] += 1; by[key][1] += float(r[vi].replace(",", "")) / 1e3
tot = sum(v[1] for v in by.values())
print(f"total {tot/1e3:.3f} ms over {sum(v[0] for v in by.values())} launches")
for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}%  n={n:5d}  avg={t/n:9.1f} us  {' '.join(k)}")
