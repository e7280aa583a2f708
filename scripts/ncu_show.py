"""ncu_show.py REPORT [kernel-substring ...]: key metrics + stall mix of the kernels in an .ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]; pats = sys.argv[2:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines())); hdr = r[0]; ki = hdr.index("Kernel Name")
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_ld.sum",
        "smsp__inst_executed_op_global_st.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum"]
seen = set()
for row in r[2:]:
    k = row[ki].split("(")[0]
    if k in seen or (pats and not any(p in k for p in pats)):
        continue
    seen.add(k); print(k)
    for w in want:
        if w in hdr: print("   ", w, row[hdr.index(w)], r[1][hdr.index(w)])
    st = [(float(row[i].replace(",", "")), hdr[i]) for i in range(len(hdr))
          if "smsp__pcsamp_warps_issue_stalled" in hdr[i] and "not_issued" not in hdr[i] and row[i].replace(",", "").replace(".", "").isdigit()]
    tot = sum(v for v, _ in st) or 1
    print("    stalls:", ", ".join(f"{n.split('stalled_')[1]} {100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:8]))
