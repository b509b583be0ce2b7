"""Condense `ncu -i X.ncu-rep --page raw --csv` exports into one small csv (the per-launch figures profiles/ keeps):
   python tools/ncu_summary.py "capture description" raw1.csv [raw2.csv ...] > profiles/rNN_ncu_<kernel>.csv"""
import csv, sys
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
w = csv.writer(sys.stdout)
desc = sys.argv[1]
first = True
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) if k in hdr else -1 for k in KEYS]
    if first:
        w.writerow(["capture"] + KEYS)
        w.writerow(["units"] + [units[i] if i >= 0 else "" for i in idx])
        first = False
    for r in rows[2:]:
        if len(r) == len(hdr):
            w.writerow([desc + " [" + path.split("/")[-1] + "]"] + [r[i] if i >= 0 else "" for i in idx])
