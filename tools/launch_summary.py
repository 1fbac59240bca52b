"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.  usage: launch_summary.py file.csv"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except Exception:
        continue
    v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    a = agg.setdefault(r[ki][:64], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in agg.values())
for n, (c, v) in agg.items():
    print("%-66s %4d launches %10.1f us total %9.1f us each %6.2f%%" % (n, c, v, v / c, 100 * v / tot))
