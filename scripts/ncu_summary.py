"""Key metrics of one kernel of an ncu raw page (csv): python scripts/ncu_summary.py <raw.csv> [kernel-substring]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
sub = sys.argv[2] if len(sys.argv) > 2 else "k_solve"
iK = hdr.index("Kernel Name")
data = [r for r in rows[2:] if sub in r[iK]]
want = ["Kernel Name", "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__icc_request_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for r in data:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(w, units[i], r[i])
    print()
