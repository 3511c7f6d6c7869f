"""Aggregate an ncu launch-list CSV (tools/one_step.py) per kernel for the LAST training step: python tools/agg_launches.py launches.csv [top_n]"""
import csv, collections, sys
f = sys.argv[1]
rows = [r for r in csv.reader(open(f)) if len(r) > 10]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
seq = [(r[ix['Kernel Name']], float(r[ix['Metric Value']].replace(',', ''))) for r in rows[1:] if r[ix['Metric Name']] == 'gpu__time_duration.sum']
last = max(i for i, (k, _) in enumerate(seq) if 'pose_fwd' in k)
step = seq[last:]
print(len(step), "launches in the last step,", round(sum(t for _, t in step) / 1e3, 1), "us")
agg = collections.OrderedDict()
for k, t in step:
    k = k[:70]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(t for _, t in step)
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{k:72s} {n:4d} {t/1e3:9.1f} us {100*t/tot:5.1f}%")
if len(sys.argv) > 3:
    for k, t in step:
        if 'fused' in k or 'chain' in k: print(k[:60], round(t / 1e3, 1))
