import csv, subprocess, sys
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
h=rows[0]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','launch__grid_size','lts__t_bytes.sum','l1tex__t_bytes.sum','sm__inst_executed.sum','smsp__inst_executed.avg.per_cycle_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active','launch__shared_mem_per_block_dynamic','sm__cycles_elapsed.avg','smsp__cycles_active.avg']
extra=[x for x in h if 'pipe_tensor' in x and x not in want]
want+=extra
for r in rows[2:]:
    for w in want:
        if w in h: print(f"{w} = {r[h.index(w)][:80]} {rows[1][h.index(w)]}")
    st=[(float(r[i]),h[i]) for i in range(len(h)) if h[i].startswith('smsp__average_warps_issue_stalled') and h[i].endswith('per_issue_active.ratio') and r[i]]
    for v,n in sorted(st,reverse=True)[:6]: print(f"   stall {n[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {v:.2f}")
