"""Prints the metrics we care about from an `ncu --page raw --csv` export (one block per profiled launch)."""
import csv
import sys

WANT = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio' ]
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
extra = sys.argv[2:]
for r in rows[2:]:
    print('-----')
    for w in WANT + extra:
        if w in h:
            print(f"{w:80s} {r[h.index(w)]}")
    for i, name in enumerate(h):
        if 'warp_issue_stalled' in name and name.endswith('_per_warp_active.pct'):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 3:
                print(f"   stall {name[len('smsp__warp_issue_stalled_'):-len('_per_warp_active.pct')]:40s} {v:.1f}")
