"""Summarise an ncu launch-list CSV (gpu__time_duration.sum): per kernel name count / total / share.
usage: python tools/ncu_launches.py launches.csv [skip_first_n]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = None
agg = collections.OrderedDict()
n = 0
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    if d.get("Metric Unit", "") in ("ns", "nsecond"):
        v /= 1e3
    elif d.get("Metric Unit", "") in ("ms", "msecond"):
        v *= 1e3
    n += 1
    if n <= skip:
        continue
    name = d["Kernel Name"]
    name = name.split("(")[0][:100]
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{n - skip} launches, total {tot / 1e3:.2f} ms (cold-cache, serialised: compare SHARES)")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{len(v):5d} x {sum(v) / len(v):9.1f} us  = {sum(v) / 1e3:8.3f} ms  {100 * sum(v) / tot:5.1f}%  {k}")
