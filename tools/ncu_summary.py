"""Print the metrics that matter from an .ncu-rep (development aid): python tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed_op_branch.sum', 'sm__inst_executed_pipe_lsu', 'smsp__pcsamp_warps_issue_stalled']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    print(d['Kernel Name'])
    for k in hdr:
        if any(s in k for s in KEYS) and d[k] not in ('', '0'):
            print('   %-90s %s %s' % (k, d[k], u[k]))
