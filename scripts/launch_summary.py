"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel instantiation for one step.

    python scripts/launch_summary.py gpurun_out/launches.csv [step_index]

A step starts at the first kernel of the forward pass (normalize_u8 / patch_embed_fwd / patchify)."""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        h = r; start = i; break
ci = {k: j for j, k in enumerate(h)}
L = []
for r in rows[start + 1:]:
    if len(r) < len(h): continue
    v = float(r[ci['Metric Value']].replace(',', '')); unit = r[ci['Metric Unit']]
    v = v / 1000 if unit == 'ns' else v * 1000 if unit == 'ms' else v
    L.append((r[ci['Kernel Name']], r[ci['Grid Size']], r[ci['Block Size']], v))
first = ('normalize_u8', 'patchify', 'patch_embed_fwd')   # first kernel of a step's forward pass
idx = [i for i, l in enumerate(L) if any(f in l[0] for f in first)]
idx = [i for j, i in enumerate(idx) if j == 0 or i - idx[j - 1] > 3]   # normalize_u8 is followed by patch_embed_fwd
k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
step = L[idx[k]:idx[k + 1]] if k + 1 < len(idx) else L[idx[k]:]
tot = sum(l[3] for l in step)
print(f'step {k}: launches {len(step)}, sum of kernel time {tot:.1f} us')
agg = collections.OrderedDict()
for n, g, b, t in step:
    n = re.sub(r'\(.*$', '', n).replace('void ', '').replace('vitk::', '')
    if n.startswith('at::'): n = 'at:: ' + re.sub(r'<.*', '', n[4:])[:60]
    a = agg.setdefault(n[:120], [0, 0.0]); a[0] += 1; a[1] += t
print('| kernel | launches | us | avg us | share |\n|---|---|---|---|---|')
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| `{n}` | {c} | {t:.1f} | {t/c:.1f} | {100*t/tot:.1f}% |')
