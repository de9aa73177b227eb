"""Development aid: top SASS instructions of an ncu report by a stall column (default stall_long_sb)."""
import csv, subprocess, sys
rep = sys.argv[1]; col = sys.argv[2] if len(sys.argv) > 2 else 'stall_long_sb'; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None; ln = None; src = ''; items = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; idx = {k: i for i, k in enumerate(hdr)}; continue
    if hdr is None or len(r) < 10: continue
    if r[0] != '':
        try: ln = int(r[0]); src = r[1][:60]
        except ValueError: pass
    elif r[2] != '...':
        try: v = float(r[idx[col]])
        except ValueError: v = 0
        items.append((v, cur, ln, r[3][:60], src))
tot = sum(i[0] for i in items)
print('total', col, tot)
for v, f, l, sass, s in sorted(items, key=lambda x: -x[0])[:top]:
    print(f'{v:7.0f} {v/tot*100:5.1f}% {f}:{l} {sass:60s} | {s}')
