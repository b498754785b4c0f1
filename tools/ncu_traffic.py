"""DRAM traffic per launch of the step's kernels from an `ncu --set full` report -> profiles/kernel_traffic.json, the
file bench.py reads `roofline.traffic` from (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the captured
launches of each kernel).

    python tools/ncu_traffic.py gpurun_out/step.ncu-rep [more.ncu-rep ...]  [--out profiles/kernel_traffic.json]
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

TAGS = [("gemm_as_kernel<carel::EpiBowFused", "bow_passF"), ("ygrad_kernel<1>", "bow_dw"), ("ygrad_kernel<(bool)1>", "bow_dw"),
        ("ygrad_kernel<0>", "bow_dz"), ("ygrad_kernel<(bool)0>", "bow_dz"), ("gemm_as_kernel<carel::EpiMMD", "mmd_gram_fwd"),
        ("PolMmdBwd", "mmd_gram_bwd"), ("bow_pack_ell_kernel", "bow_pack_ell"), ("gemm_2sm_kernel<carel::EpiMMD", "mmd_gram_fwd_d768"),
        ("EpiHsic", "hsic_gram_fwd"), ("PolHsicBwd", "hsic_gram_bwd"), ("EpiBowFused", "bow_passF"), ("EpiMMD", "mmd_gram_fwd")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    out_path = Path("profiles/kernel_traffic.json")
    if "--out" in sys.argv:
        out_path = Path(sys.argv[sys.argv.index("--out") + 1])
        args = [a for a in args if a != str(out_path)]
    acc = {}
    for rep in args:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        ci = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[ci["Kernel Name"]]
            tag = next((t for pat, t in TAGS if pat in name), None)
            if tag is None:
                continue
            tot = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[ci[m]]) * UNIT.get(units[ci[m]], 1.0)
            us = float(r[ci["gpu__time_duration.sum"]]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(units[ci["gpu__time_duration.sum"]], 1.0)
            acc.setdefault(tag, []).append((tot, us, Path(rep).name))
    res = json.loads(out_path.read_text()) if out_path.exists() else {}
    meta = res.setdefault("_source", {})
    for tag, v in acc.items():
        res[tag] = round(sum(t for t, _, _ in v) / len(v))
        meta[tag] = {"report": v[0][2], "launches": len(v), "ncu_us_per_launch": round(sum(u for _, u, _ in v) / len(v), 1)}
    out_path.write_text(json.dumps(res, indent=1, sort_keys=True) + "\n")
    print(json.dumps(res, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
