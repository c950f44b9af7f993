"""Text summary of an .ncu-rep (run here, no GPU needed):  python tools/ncu_report.py gpurun_out/prof.ncu-rep > profiles/x.txt"""
import csv, io, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("kernel:", r[idx["Kernel Name"]][:110])
    print("  grid", r[idx.get("launch__grid_size", 0)], "block", r[idx.get("launch__block_size", 0)])
    for w in WANT:
        if w in idx:
            print(f"  {w:70s} {r[idx[w]]:>16s} {units[idx[w]]}")
