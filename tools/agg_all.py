"""Aggregate the LAST iteration of an ncu launch-list CSV, iteration = launches since the last occurrence of a marker kernel:
python tools/agg_all.py launches.csv <marker substring> [top_n]"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
seq = [(r[ix['Kernel Name']], float(r[ix['Metric Value']].replace(',', ''))) for r in rows[1:] if r[ix['Metric Name']] == 'gpu__time_duration.sum']
marker = sys.argv[2]
last = max(i for i, (k, _) in enumerate(seq) if marker in k)
step = seq[last:]
tot = sum(t for _, t in step)
print(len(step), "launches in the last iteration,", round(tot / 1e3, 1), "us")
agg = collections.OrderedDict()
for k, t in step:
    a = agg.setdefault(k[:78], [0, 0.0]); a[0] += 1; a[1] += t
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print(f"{k:80s} {n:4d} {t/1e3:9.1f} us {100*t/tot:5.1f}%")
