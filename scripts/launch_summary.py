"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel instantiation for one step.

    python scripts/launch_summary.py gpurun_out/launches.csv [step_index]

A step starts at each patchify_kernel launch (first kernel of the forward pass)."""
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
idx = [i for i, l in enumerate(L) if 'patchify' in l[0] or 'patch_embed' in l[0]]
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
