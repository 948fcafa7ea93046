"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    v *= scale
    t = tot[name]
    t[0] += 1; t[1] += v; t[2] = max(t[2], v)
total = sum(t[1] for t in tot.values())
print(f"{'kernel':40s} {'launches':>8s} {'total_us':>12s} {'share':>7s} {'avg_us':>10s} {'max_us':>10s}")
for name, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:40s} {t[0]:8d} {t[1]:12.1f} {100*t[1]/total:6.1f}% {t[1]/t[0]:10.2f} {t[2]:10.2f}")
print(f"{'TOTAL':40s} {sum(t[0] for t in tot.values()):8d} {total:12.1f}")
