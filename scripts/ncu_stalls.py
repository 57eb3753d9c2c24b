"""Aggregate warp-stall samples of an `ncu --page source --csv` dump by stall reason and by opcode.
python scripts/ncu_stalls.py file.csv [first_line last_line]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
by_reason = collections.Counter(); by_op = collections.Counter(); ex_op = collections.Counter()
tot = 0
for k, r in enumerate(rows[2:]):
    if not (lo <= k <= hi): continue
    try: n = float(r[ci['# Samples']])
    except Exception: continue
    src = r[ci['Source']].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')
    op = op.split('.')[0]
    if op in ('EXIT',) or 'SYNCS' in src or op == 'NANOSLEEP': continue   # idle / barrier waits
    tot += n
    by_op[op] += n
    ex_op[op] += float(r[ci['Instructions Executed']] or 0)
    for s in stalls:
        by_reason[s] += float(r[ci[s]] or 0)
print('samples (excluding EXIT / mbarrier waits):', tot)
for s, v in by_reason.most_common(10): print(f'  {s:28s} {v:8.0f} {100*v/tot:5.1f}%')
print('by opcode: samples, warp-instructions executed')
for o, v in by_op.most_common(25): print(f'  {o:12s} {v:8.0f} {100*v/tot:5.1f}%   exec {ex_op[o]:12.0f}')
print('total warp-instructions', sum(ex_op.values()))
