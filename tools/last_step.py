"""Per-kernel breakdown of the LAST step in an ncu launch list: python tools/last_step.py file.csv {maskgit|encode|decode|train}"""
import csv, re, sys
fn, kind = sys.argv[1], sys.argv[2]
lines = [l for l in open(fn) if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
names = [re.sub(r'\(.*', '', r['Kernel Name']).replace('void phk::', '').replace('phk::', '') for r in rows]
key = {'maskgit': 'token_embed_kernel', 'train': 'token_embed_kernel', 'decode': 'lfq_codes_kernel'}.get(kind, 'patchify_ln')
idxs = [i for i, nm in enumerate(names) if key in nm]
start = idxs[-2] if kind == 'encode' else idxs[-1]
agg, tot = {}, 0.0
for r, nm in list(zip(rows, names))[start:]:
    if 'at::' in nm:
        continue
    v = float(r['Metric Value'].replace(',', '')) / 1000
    k = (nm[:44], r['Grid Size'])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(fn, 'last step: sum of kernel durations', round(tot, 1), 'us')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30 if kind == 'train' else 18]:
    print(f"  {a[1]/tot*100:5.1f}% {k[0]:44s} grid={k[1]:16s} n={a[0]:3d} avg={a[1]/a[0]:8.1f} us")
