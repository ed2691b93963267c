"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share)."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    k = row["Kernel Name"].split("(")[0][:70]
    try: v = float(row["Metric Value"].replace(",", ""))
    except ValueError: continue
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':70s} {'n':>5s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} {v[0]:5d} {v[1]/1e6:10.3f} {v[1]/v[0]/1e3:10.1f} {v[1]/tot*100:6.1f}%")
