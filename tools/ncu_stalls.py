"""Development aid: per-source-line stall breakdown (long scoreboard / no instruction ...) of an ncu report."""
import csv, subprocess, sys
rep = sys.argv[1]; key = sys.argv[2] if len(sys.argv) > 2 else 'stall_long_sb'; top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, agg = None, None, {}
def f(x):
    try: return float(x)
    except Exception: return 0.0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; ki = hdr.index(key); continue
    if hdr is None or len(r) <= ki or r[0] == '': continue
    try: ln = int(r[0])
    except ValueError: continue
    k = (cur, ln)
    a = agg.setdefault(k, [r[1][:110], 0.0, 0.0])
    a[1] += f(r[ki]); a[2] += f(r[6])
tot = sum(a[1] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print(key, 'total', tot, 'of', ts, 'samples')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{k[0]}:{k[1]:4d} {a[1]/max(tot,1)*100:5.1f}% | {a[0]}')
