"""Hot SASS of kernel number k (0-based) in an ncu report: python tools/ncu_src.py rep.ncu-rep k [min_pct]"""
import csv, io, subprocess, sys
rep, k = sys.argv[1], int(sys.argv[2])
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
b = blocks[k]
hdr = b["rows"][0]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in b["rows"][1:] if len(r) == len(hdr)]
tot = sum(float(r[ci["# Samples"]] or 0) for r in data)
print(b["name"][:120], "total samples", tot)
for idx, r in enumerate(data):
    s = float(r[ci["# Samples"]] or 0)
    if s >= tot * min_pct / 100:
        print(f"#{idx:5d} {100 * s / tot:5.1f}% ex={float(r[ci['Instructions Executed']] or 0):9.0f}  {r[ci['Source']].strip()[:110]}")
