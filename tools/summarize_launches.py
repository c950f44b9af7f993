"""Summarise an `ncu --csv --metrics gpu__time_duration.sum` launch list: time per kernel name."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = {n: i for i, n in enumerate(r)}
        continue
    if len(r) < len(hdr) or r[hdr["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[hdr["Kernel Name"]])[:70]
    v = float(r[hdr["Metric Value"]].replace(",", ""))
    unit = r[hdr["Metric Unit"]]
    v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)  # -> us
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot/1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}%  x{n:<5d} {k}")
