"""Development aid: per-source-line summary of an ncu report's source page (samples, instructions, threads/instr)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, agg = None, None, []
def f(x):
    try: return float(x)
    except Exception: return 0.0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < 10 or r[0] == '': continue
    try: ln = int(r[0])
    except ValueError: continue
    agg.append((cur, ln, r[1][:100], f(r[6]), f(r[7]), f(r[8])))
ts, ti = sum(a[3] for a in agg), sum(a[4] for a in agg)
print('total samples', ts, 'warp-instructions', ti)
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    print(f'{a[0]}:{a[1]:4d} samp {a[3]/ts*100:5.1f}% inst {a[4]/ti*100:5.1f}% thr/inst {a[5]/max(a[4],1):4.1f} | {a[2]}')
if len(sys.argv) > 3:      # line-range buckets "a-b,c-d" of mfg_obs.cu
    for rng in sys.argv[3].split(','):
        lo, hi = map(int, rng.split('-'))
        sel = [a for a in agg if a[0] == 'mfg_obs.cu' and lo <= a[1] <= hi]
        print(f'lines {lo}-{hi}: samp {sum(a[3] for a in sel)/ts*100:5.1f}% inst {sum(a[4] for a in sel)/ti*100:5.1f}%')
    sel = [a for a in agg if a[0] != 'mfg_obs.cu']
    print(f'other files: samp {sum(a[3] for a in sel)/ts*100:5.1f}% inst {sum(a[4] for a in sel)/ti*100:5.1f}%')
