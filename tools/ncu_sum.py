"""Key metrics per kernel from an ncu raw page csv on stdin."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum']
stall = [h for h in hdr if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
ki = hdr.index('Kernel Name')
for r in rows[2:]:
    print(r[ki][:90])
    for k in keys:
        if k in hdr:
            print(f"   {k:70s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
    st = []
    for h in stall:
        try:
            st.append((float(r[hdr.index(h)]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
        except ValueError:
            pass
    print("   stalls/issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:7]))
