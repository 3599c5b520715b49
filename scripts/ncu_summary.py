"""Summarise an ncu --page raw --csv export: key metrics per kernel launch. Usage: ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__grid_size', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
stall = [h for h in hdr if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('.ratio')]
if not stall:
    stall = [h for h in hdr if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct')]
for r in rows[2:]:
    print("=" * 110)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:80s} {r[i]} {units[i]}")
    vals = []
    for h in stall:
        try:
            vals.append((float(r[hdr.index(h)].replace(',', '')), h))
        except ValueError:
            pass
    for v, h in sorted(vals, reverse=True)[:7]:
        print(f"     stall {h:90s} {v:.3f}")
