"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, sys
path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"') or l.startswith('ID,') or l[:1].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0]); total = 0
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', ''))
    if row['Metric Unit'] == 'ns': v /= 1e3
    elif row['Metric Unit'] == 'ms': v *= 1e3
    name = row['Kernel Name'].split('(')[0][:78]
    agg[name][0] += 1; agg[name][1] += v; total += v
print(f"total {total/1e3:.2f} ms over {sum(c for c,_ in agg.values())} launches")
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"| `{k}` | {c} | {t:.0f} | {100*t/total:.1f}% | {t/c:.1f} |")
