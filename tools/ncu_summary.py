"""Condense an .ncu-rep (ncu -i ... --page raw --csv) into the few numbers the
design notes quote.  usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    print("kernel:", name[:100])
    for k in KEYS:
        if k in h:
            print(f"  {k:70s} {r[h.index(k)]:>18s} {units[h.index(k)]}")
    print("  pipes (% of peak, active):")
    for i, c in enumerate(h):
        if (c.startswith("sm__inst_executed_pipe_") or c.startswith("sm__pipe_")) and "pct_of_peak_sustained_active" in c:
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if v >= 1.0:
                print(f"    {c:76s} {v:6.1f}")
    st = [(float(r[i].replace(",", "")), c) for i, c in enumerate(h)
          if c.startswith("smsp__pcsamp_warps_issue_stalled_") and not c.endswith("_not_issued")]
    tot = sum(v for v, _ in st) or 1
    print("  warp stall samples (share):")
    for v, c in sorted(st, reverse=True)[:8]:
        print(f"    {c.replace('smsp__pcsamp_warps_issue_stalled_', ''):28s} {100 * v / tot:5.1f} %")
