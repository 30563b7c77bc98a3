"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_summary.py file.csv [last_n]"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
seq = [(x["Kernel Name"].split("(")[0][-34:], x["Grid Size"], float(x["Metric Value"].replace(",", ""))) for x in rows]
n = int(sys.argv[2]) if len(sys.argv) > 2 else len(seq)
last = seq[-n:]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, g, t in last:
    agg[k][0] += 1; agg[k][1] += t
tot = sum(t for _, _, t in last)
print(f"{len(last)} launches, {tot/1e3:.1f} us")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:36s} x{c:4d} {t/1e3:10.1f} us  {100*t/tot:5.1f}%")
if "-v" in sys.argv:
    for k, g, t in last:
        print(f"{k:36s} {g:>16s} {t/1e3:9.1f}")
