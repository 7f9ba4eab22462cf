"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr = r; start = i + 1; break
ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
d = collections.defaultdict(list)
for r in rows[start:]:
    if len(r) > vi:
        v = float(r[vi].replace(",", ""))
        if r[ui] == "ns": v /= 1000.0
        elif r[ui] == "ms": v *= 1000.0
        d[r[ki].split("(")[0]].append(v)
tot = sum(sum(v) for v in d.values())
print(f"{'kernel':40s} {'n':>5s} {'avg us':>10s} {'total us':>11s} {'share':>7s}")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:40s} {len(v):5d} {sum(v)/len(v):10.2f} {sum(v):11.1f} {sum(v)/tot:7.3f}")
print(f"{'TOTAL':40s} {sum(len(v) for v in d.values()):5d} {'':10s} {tot:11.1f}")
