"""profiles/r02_stage_ncu.md from the raw page export of `ncu --set full ... -k regex:aio_stage_kernel -c 8 python scripts/prof_stage.py`
(`ncu -i X.ncu-rep --page raw --csv > raw.csv`).  Launches 0 / 2 / 4 / 6 = eval stage 1, train stage 1, eval stage 2, train stage 2.
Usage: python scripts/summarize_stage_ncu.py raw.csv > profiles/r02_stage_ncu.md"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
pick = [data[i] for i in (0, 2, 4, 6)]
M = [('duration', 'gpu__time_duration.sum'), ('grid', 'launch__grid_size'), ('block', 'launch__block_size'), ('regs / thread', 'launch__registers_per_thread'),
     ('tensor pipe active % (elapsed)', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'),
     ('tensor pipe active % (while the SM is active)', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
     ('tcgen05.ld executed (warp instructions)', 'smsp__sass_inst_executed_op_tmem_ldt.sum'),
     ('shared-memory pipe active % (elapsed)', 'sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_elapsed'),
     ('issue slots % (elapsed)', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'),
     ('warps active % of peak', 'sm__warps_active.avg.pct_of_peak_sustained_active'),
     ('DRAM read', 'dram__bytes_read.sum'), ('DRAM write', 'dram__bytes_write.sum'),
     ('L2 sectors % of peak', 'lts__t_sectors.avg.pct_of_peak_sustained_elapsed'),
     ('smem bank conflicts ld', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum'), ('smem bank conflicts st', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum'),
     ('stall long scoreboard / issue', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'),
     ('stall branch resolving / issue', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio'),
     ('stall barrier / issue', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio'),
     ('stall wait / issue', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio'),
     ('stall short scoreboard / issue', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio'),
     ('stall math pipe throttle / issue', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio')]
print('| metric | eval, stage 1 (C=12, 16x16, width 32), 24 blocks, B=2048 | train, stage 1, B=128 | eval, stage 2 (C=48, 8x8, width 64), B=2048 | train, stage 2, B=128 |')
print('|---|---:|---:|---:|---:|')
for name, key in M:
    if key not in hdr:
        continue
    i = hdr.index(key)
    print('| %s | %s |' % (name, ' | '.join('%s %s' % (r[i], units[i]) for r in pick)))
