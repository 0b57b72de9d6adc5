"""Sums an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
    name = r[ki].split("(")[0]
    tot[name][0] += 1
    tot[name][1] += v
s = sum(v[1] for v in tot.values())
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {100 * t / s:5.1f} %  x{n:<4d} {k}")
print(f"{s:10.1f} us total")
