"""Writes profiles/<tag>_ncu_full_summary.csv and the traffic entries of profiles/obs_kernel_traffic.json (stamped with the
sha256 of the kernel sources) from the ncu reports gpurun_out/<dir>/prof_{obs_faithful,obs_identity,step}.ncu-rep
(ncu --set full --clock-control none, one launch each, steady-state step 200, 1 M envs).
    python tools/summarize_profiles.py r2_final r2z"""
import csv
import json
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r2_final'
DIR = sys.argv[2] if len(sys.argv) > 2 else 'r2z'
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'smsp__issue_active.avg.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
rows_out = [['report', 'kernel', 'metric', 'unit', 'value']]
traffic = {}
for name, rep in (('obs_faithful', 'prof_obs_faithful'), ('obs_identity', 'prof_obs_identity'), ('step_faithful', 'prof_step')):
    out = subprocess.run(['ncu', '-i', str(ROOT / 'gpurun_out' / DIR / f'{rep}.ncu-rep'), '--page', 'raw', '--csv'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    kn = v[h.index('Kernel Name')][:70]
    d = {}
    for i, k in enumerate(h):
        if k in KEYS or ('issue_stalled' in k and 'per_issue_active' in k):
            rows_out.append([name, kn, k, u[i], v[i]])
            d[k] = (v[i], u[i])
    traffic[name] = sum(float(d[k][0].replace(',', '')) * UNIT[d[k][1]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
csv.writer(open(ROOT / 'profiles' / f'{TAG}_ncu_full_summary.csv', 'w')).writerows(rows_out)
tp = ROOT / 'profiles' / 'obs_kernel_traffic.json'
tj = json.loads(tp.read_text())
tj['cfg4:faithful:1048576'] = traffic['obs_faithful']
tj['cfg4:identity:1048576'] = traffic['obs_identity']
from bench import kernel_source_sha
tj['source_sha256'] = kernel_source_sha()
tj['captured'] = f'{TAG}, ' + time.strftime('%Y-%m-%d')
tp.write_text(json.dumps(tj, indent=1))
print(traffic)
