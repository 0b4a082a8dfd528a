"""Offline: summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name (count, total, share)."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
iN, iV, iM = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
iU = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= iV or r[iM] != "gpu__time_duration.sum":
        continue
    v = float(r[iV].replace(",", ""))
    v = v / 1e3 if r[iU] in ("ns", "nsecond") else v
    name = re.sub(r"\(.*", "", r[iN])
    name = re.sub(r"<.*", "", name)[:70]
    c, t = agg.get(name, (0, 0.0))
    agg[name] = (c + 1, t + v)
tot = sum(t for _, t in agg.values())
print(f"{sum(c for c, _ in agg.values())} launches, {tot / 1e3:.3f} ms of kernel time (serialised, cold-cache: compare shares)")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / 1e3:8.3f} ms {100 * t / tot:5.1f}%  n={c:4d}  avg {t / c:8.1f} us  {name}")
