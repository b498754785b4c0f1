"""Key metrics of every kernel in an ncu report: python tools/ncu_key.py X.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=" * 100)
    for k in want:
        if k in d:
            print(f"{k:90s} {d[k]}  {rows[1][hdr.index(k)]}")
    st = sorted(((float(d[s] or 0), s.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for s in stalls), reverse=True)
    print("stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in st[:9]))
