"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count/total/first durations."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    name = r[ki].replace("<unnamed>::", "").split("(")[0][:48]
    agg.setdefault(name, []).append(float(r[vi].replace(",", "")) / 1e3)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':48s} {'n':>5s} {'total_us':>10s} {'share':>6s}  first durations (us)")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    big = [x for x in v if x > 8.0]
    print(f"{k:48s} {len(v):5d} {sum(v):10.1f} {100*sum(v)/tot:5.1f}%  " + " ".join(f"{x:.0f}" for x in (big[:14] if big else v[:6])))
