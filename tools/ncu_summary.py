"""Condense `ncu -i report.ncu-rep --page raw --csv` into the per-kernel lines kept under profiles/ (DRAM bytes, duration, ...)."""
import csv, sys
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        # the shared-memory data pipe (what bounds the fp32 kernels): LSU wavefronts (LDS/STS of the converters and epilogues),
        # tensor-core operand fetches, bank-conflict replays; TMA landing / store reads are not in these counters
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum", "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_red.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units = rows[hdr], rows[hdr + 1]
col = {n: i for i, n in enumerate(names)}
tot_r = tot_w = 0.0
for r in rows[hdr + 2:]:
    if len(r) < len(names): continue
    print("----")
    print("%-75s %s" % ("Kernel Name", r[col["Kernel Name"]][:70]))
    for k in KEEP:
        if k in col:
            print("%-75s %18s %s" % (k, r[col[k]], units[col[k]]))
    def mb(k):
        v, u = float(r[col[k]].replace(",", "")), units[col[k]].lower()
        return v * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1.0)
    tot_r += mb("dram__bytes_read.sum"); tot_w += mb("dram__bytes_write.sum")
print("----")
print("total dram read %.1f MB, write %.1f MB over the listed launches" % (tot_r, tot_w))
