"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]` launch list:
time (and DRAM traffic) per kernel name; optional JSON with the per-step totals of the dominant kernel.

    python tools/summarize_launches.py gpurun_out/launches.csv [profiles/r1_gemm_traffic.json]
"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
per_launch = collections.OrderedDict()   # id -> dict(name, us, rd, wr)
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = {n: i for i, n in enumerate(r)}
        continue
    if len(r) < len(hdr):
        continue
    lid = r[hdr["ID"]]
    d = per_launch.setdefault(lid, {"name": re.sub(r"\(.*", "", r[hdr["Kernel Name"]])[:70], "us": 0.0, "rd": 0.0, "wr": 0.0})
    m, unit = r[hdr["Metric Name"]], r[hdr["Metric Unit"]]
    v = float(r[hdr["Metric Value"]].replace(",", ""))
    if m == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d["rd" if "read" in m else "wr"] = v * scale
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in per_launch.values():
    a = agg[d["name"]]
    a[0] += 1; a[1] += d["us"]; a[2] += d["rd"]; a[3] += d["wr"]
tot = sum(v[1] for v in agg.values())
print(f"total {tot/1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches (ncu: cold-cache, serialised -- compare shares)")
print(f"{'ms':>9s} {'share':>6s}  {'n':>5s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s}  kernel")
for k, (n, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{t/1e3:9.3f} {100*t/tot:5.1f}%  x{n:<5d} {rd/1e6:10.1f} {wr/1e6:11.1f}  {k}")
if len(sys.argv) > 2:
    g = [(n, t, rd, wr) for k, (n, t, rd, wr) in agg.items() if "tris_umma_gemm_kernel" in k]
    out = {"kernel": "tris_umma_gemm_kernel (all instantiations)", "launches_per_step": sum(x[0] for x in g),
           "ncu_ms_per_step": sum(x[1] for x in g) / 1e3, "share_of_step": sum(x[1] for x in g) / tot,
           "dram_read_bytes_per_step": sum(x[2] for x in g), "dram_write_bytes_per_step": sum(x[3] for x in g),
           "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over "
                     "tools/profile_step.py (one eager step, batch 48)"}
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps(out))
