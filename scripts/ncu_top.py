"""Print the top stall lines of an `ncu --page source --csv` dump: python scripts/ncu_top.py file.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for k, r in enumerate(rows[2:]):
    try:
        data.append((float(r[ci['# Samples']]), k, r))
    except Exception:
        pass
tot = sum(v for v, _, _ in data)
print(rows[0][1][:90], 'total samples', tot)
for v, k, r in sorted(data, key=lambda t: -t[0])[:n]:
    st = sorted(((float(r[ci[s]]), s) for s in stalls), reverse=True)[:2]
    print(f'{v:7.0f} {100*v/tot:5.1f}% L{k:4d} {r[ci["Source"]].strip()[:60]:60s} {st[0][1]}={st[0][0]:.0f} {st[1][1]}={st[1][0]:.0f}')
