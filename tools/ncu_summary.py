"""Key metrics of an .ncu-rep (from `ncu --set full`) as a markdown table.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/x.md"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct"]
idx = [hdr.index(w) for w in want if w in hdr]
print("| " + " | ".join(hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in idx) + " |")
print("|" + "---|" * len(idx))
for r in rows[2:]:
    print("| " + " | ".join(r[i].split("(")[0][-40:] if hdr[i] == "Kernel Name" else r[i] for i in idx) + " |")
