"""One row per `ncu --set full` capture (raw page) with the metrics the round summary cites.
usage: python tools/ncu_kernels.py out.csv label=report.ncu-rep ..."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
out = csv.writer(open(sys.argv[1], "w"))
out.writerow(["capture", "kernel"] + KEYS)
for arg in sys.argv[2:]:
    label, rep = arg.split("=", 1)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h = rows[0]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        out.writerow([label, name[:60]] + [(r[h.index(k)] + " " + rows[1][h.index(k)]).strip() if k in h else "" for k in KEYS])
