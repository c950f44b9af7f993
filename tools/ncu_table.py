"""One line per launch of an .ncu-rep (run here, no GPU needed):  python tools/ncu_table.py x.ncu-rep > profiles/x.txt
columns: duration, tensor-pipe active %, DRAM bytes read+written, DRAM / L2 / SM throughput % of peak, grid, kernel."""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
def f(r, k, d=0.0):
    try:
        return float(r[idx[k]].replace(",", ""))
    except Exception:
        return d
def scale(k, v):
    u = units[idx[k]] if k in idx else ""
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3 * 1e3, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)
print(f"{'#':>3} {'us':>8} {'tensor%':>8} {'dramMB':>8} {'dram%':>6} {'l2%':>6} {'sm%':>6} {'grid':>6} {'regs':>5}  kernel")
tot = 0.0
for i, r in enumerate(rows[2:]):
    us = scale("gpu__time_duration.sum", f(r, "gpu__time_duration.sum"))
    tot += us
    mb = (scale("dram__bytes_read.sum", f(r, "dram__bytes_read.sum")) + scale("dram__bytes_write.sum", f(r, "dram__bytes_write.sum"))) / 1e6
    name = r[idx["Kernel Name"]]
    print(f"{i:3d} {us:8.2f} {f(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):8.1f} {mb:8.2f} "
          f"{f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} {f(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{f(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} {r[idx['launch__grid_size']]:>6} {r[idx['launch__registers_per_thread']]:>5}  {name[:100]}")
print(f"total {tot:.1f} us over {len(rows) - 2} launches (cold-cache, serialised ncu replays)")
