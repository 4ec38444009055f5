"""Prints selected metrics of an ncu raw CSV page (ncu -i x.ncu-rep --page raw --csv), one row per metric."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_global_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "sm__inst_executed.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.per_cycle_active", "launch__grid_size", "launch__block_size"]
want += [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") or "warp_issue_stalled" in h and "ratio" in h]
seen = set()
for w in want:
    if w in hdr and w not in seen:
        seen.add(w)
        i = hdr.index(w)
        print(w, [r[i][:28] for r in rows[2:]])
