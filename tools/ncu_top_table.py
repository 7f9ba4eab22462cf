"""Selected `ncu --set full` metrics of every launch in an .ncu-rep as one CSV row per launch (profiles/*_full.csv).
usage: ncu_top_table.py rep out.csv"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
want = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
ki = h.index("Kernel Name")
cols = [h.index(w) for w in want if w in h]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["Kernel Name"] + [h[c] for c in cols])
    w.writerow([""] + [units[c] for c in cols])
    for r in rows[2:]:
        w.writerow([r[ki].split("(")[0]] + [r[c] for c in cols])
