"""Summarise an ncu report (.ncu-rep) into a small text file for profiles/: key raw metrics of
each profiled kernel + per-source-file instruction/thread-utilisation/stall-sample breakdown."""
import csv, subprocess, sys, io
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct", "smsp__inst_executed.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
with open(out, "w") as f:
    f.write(f"# ncu summary of {rep}\n")
    for r in rows[2:]:
        f.write("\n")
        for h, u, v in zip(hdr, units, r):
            if h in KEYS or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                f.write(f"{h} [{u}] = {v}\n")
print(open(out).read()[:3000])
