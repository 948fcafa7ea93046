"""ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch, --csv) ->
profiles/roofline_traffic.json: DRAM bytes of the mf_factor_* launches of ONE Newton iteration, plus a per-kernel
table on stdout. usage: make_traffic.py launches.csv scenarios iterations_in_the_capture"""
import collections, csv, json, sys

path, S, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
with open(path) as fh:
    lines = [l for l in fh if l.startswith('"')]
unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
launch = collections.OrderedDict()
for r in csv.DictReader(lines):
    d = launch.setdefault(r["ID"], {"k": r["Kernel Name"].split("::")[-1].split("(")[0]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * unit[r["Metric Unit"]]
tab = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in launch.values():
    t = tab[d["k"]]
    t[0] += 1
    t[1] += d["gpu__time_duration.sum"]
    t[2] += d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
total = sum(t[1] for t in tab.values())
print(f"{'kernel':36s} {'launches':>8s} {'ms':>9s} {'share':>7s} {'GB':>8s} {'GB/s':>7s}")
for k, t in sorted(tab.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:36s} {t[0]:8d} {t[1]*1e3:9.2f} {100*t[1]/total:6.1f}% {t[2]/1e9:8.2f} {t[2]/t[1]/1e9:7.0f}")
fac = [d for d in launch.values() if d["k"].startswith("mf_factor")]
out = {"scenarios": S,
       "mf_factor_kernel_dram_bytes_per_factor_phase": sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in fac) / iters,
       "factor_launches": len(fac) // iters, "ncu_sum_ms": 1e3 * sum(d["gpu__time_duration.sum"] for d in fac) / iters,
       "source": f"profiles/{path.split('/')[-1]} (ncu dram__bytes_read/write.sum summed over the factor launches of one "
                 "NR iteration)"}
with open("profiles/roofline_traffic.json", "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out))
