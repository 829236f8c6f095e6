#!/usr/bin/env python
"""Print the key metrics of an .ncu-rep (raw page) -- used to write the summaries under profiles/."""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'gpu__dram_throughput.avg.pct',
        'sm__warps_active.avg.per_cycle_active', 'launch__registers_per_thread ', 'launch__grid_size', 'launch__block_size',
        'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum ', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors.sum ', 'lts__throughput.avg.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct', 'smsp__pcsamp_warps_issue_stalled', 'smsp__pcsamp_sample_count',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor', 'smsp__warps_eligible.avg.per_cycle_active', 'launch__shared_mem_per_block_dynamic',
        'sm__cycles_elapsed.max ', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ', 'smsp__inst_executed_op_shared',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ', 'sm__inst_executed_pipe_fmaheavy', 'sm__pipe_fma_cycles_active.avg.pct',
        'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    print('==', vals[hdr.index('Kernel Name')][:60])
    for h, u, v in zip(hdr, units, vals):
        if any(k in h + ' ' for k in KEYS) and '_not_issued' not in h:
            print(f'{h:90s} {u:12s} {v}')
