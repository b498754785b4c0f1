"""Summarise an ncu source-page CSV: hottest SASS instructions by stall samples.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[2:]):
    if len(r) != len(hdr):
        continue
    s = float(r[ci["# Samples"]] or 0)
    ex = float(r[ci["Instructions Executed"]] or 0)
    st = {h: float(r[ci[h]] or 0) for h in stall_cols}
    data.append((s, idx, r[ci["Source"]].strip(), ex, st))
tot = sum(d[0] for d in data)
print(f"instructions: {len(data)}  total samples: {tot:.0f}")
agg = {}
for d in data:
    for k, v in d[4].items():
        agg[k] = agg.get(k, 0) + v
print("stall totals:", {k: int(v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for s, idx, src, ex, st in sorted(data, reverse=True)[:top]:
    main = max(st.items(), key=lambda kv: kv[1])
    print(f"{s:7.0f} {100 * s / tot:5.1f}%  #{idx:5d} exec {ex:9.0f}  {main[0]:<22} {src[:90]}")
