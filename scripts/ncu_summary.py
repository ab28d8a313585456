"""Summarise an .ncu-rep (raw page) into the handful of metrics the profiles/ notes quote.
Usage: python scripts/ncu_summary.py <file.ncu-rep> [more metric substrings...]"""
import csv, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
        "smsp__inst_executed.sum", "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
for k in KEYS + [h for h in hdr if any(e in h for e in extra)]:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k} [{units[i]}]: " + " | ".join(r[i] for r in data))
