"""One step of an ncu launch list taken with
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
in launch order: time, DRAM bytes and achieved DRAM GB/s per launch (fraction of the measured copy peak).
usage: python tools/ncu_hbm.py launches.csv [hbm_peak_GBs]"""
import csv, json, sys
peak = float(sys.argv[2]) if len(sys.argv) > 2 else json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, L, order = None, {}, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    i = int(d["ID"])
    if i not in L:
        L[i] = {"name": d["Kernel Name"].split("(")[0][:84], "stream": d.get("Stream", "")}
        order.append(i)
    v, u, m = float(d["Metric Value"].replace(",", "")), d["Metric Unit"], d["Metric Name"]
    if m == "gpu__time_duration.sum":
        L[i]["us"] = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    else:
        L[i][m] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
starts = [k for k, i in enumerate(order) if "pack_x_ones" in L[i]["name"]]
sel = order[starts[-2]:starts[-1]] if len(starts) >= 2 else order
tot = sum(L[i]["us"] for i in sel)
ours = [i for i in sel if "carel::" in L[i]["name"]]
print(f"{len(sel)} launches in the step, {tot:.1f} us serialised (cold cache); carel kernels: {len(ours)} launches, "
      f"{sum(L[i]['us'] for i in ours):.1f} us; HBM peak used for the fraction: {peak:.0f} GB/s (MEASURED_PEAKS.json)")
print(f"{'us':>8s} {'DRAM MB':>9s} {'GB/s':>7s} {'frac':>5s}  stream  kernel")
for i in sel:
    e = L[i]
    b = e.get("dram__bytes_read.sum", 0) + e.get("dram__bytes_write.sum", 0)
    gbs = b / e["us"] / 1e3 if e["us"] else 0.0
    print(f"{e['us']:8.1f} {b / 1e6:9.1f} {gbs:7.0f} {gbs / peak:5.2f}  {e['stream']:>6s}  {e['name']}")
