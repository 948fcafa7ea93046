"""Per-source-line warp-state samples and executed instructions of one kernel from
`ncu -i rep --page source --csv --print-source cuda,sass -k regex:<kernel>` (sections repeat per kernel in capture order).
usage: ncu_lines.py file.csv [section_index_of_kernel=last] [top=25]"""
import csv, sys, collections
path = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = list(csv.reader(open(path)))
# split into kernels: a kernel's dump starts at a "File Path" row whose previous kernel ended; group File sections
sections = []
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = {"file": r[1], "hdr": None, "rows": []}
        sections.append(cur)
    elif r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(r)
# kernels are separated where the same file name re-appears after other files; take groups of consecutive sections
groups, seen = [], set()
g = []
for s in sections:
    if s["file"] in seen:
        groups.append(g); g = []; seen = set()
    seen.add(s["file"]); g.append(s)
groups.append(g)
which = int(sys.argv[2]) if len(sys.argv) > 2 else len(groups) - 1
print(f"{len(groups)} kernel dumps; showing #{which}")
agg = collections.defaultdict(lambda: [0, 0, ""])
tot_s = tot_i = 0
for s in groups[which]:
    h = s["hdr"]
    i_line, i_src = 0, 1
    i_samp = h.index("# Samples")
    i_inst = h.index("Instructions Executed")
    line, src = None, ""
    for r in s["rows"]:
        if r[0] != "":
            line, src = r[0], r[1]
        try:
            sa, ins = int(r[i_samp] or 0), int(r[i_inst] or 0)
        except ValueError:
            continue
        key = (s["file"].split("/")[-1], int(line))
        agg[key][0] += sa; agg[key][1] += ins; agg[key][2] = src.strip()[:110]
        tot_s += sa; tot_i += ins
print(f"samples {tot_s}, warp instructions {tot_i}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{key[0]:14s}:{key[1]:4d} {100*v[0]/max(1,tot_s):5.1f}% samp {100*v[1]/max(1,tot_i):5.1f}% inst  {v[2]}")
