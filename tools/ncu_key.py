"""Development aid: key metrics of the first kernel in an ncu report."""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); h = rows[0]; v = rows[2]
keys = ('gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__shared_mem_per_block_dynamic')
for i, k in enumerate(h):
    if k in keys or ('issue_stalled' in k and 'per_issue_active' in k and float(v[i]) > 0.25):
        print(k, v[i], rows[1][i])
