"""Blackwell opcode histogram of every kernel in libcarel_head.so (cuobjdump -sass): the tcgen05 / TMEM / TMA evidence
the B200_PROFILING guide names (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy,
LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier, REDG = red.global, MUFU.EX2).
    python tools/sass_histogram.py [path/to/lib.so] > profiles/r02_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "carel_vae_b200/libcarel_head.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "REDG", "MUFU.EX2", "FFMA2",
        "FFMA", "HMMA", "total"]
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        c = counts[kern]
        c["total"] += 1
        base = op.split(".")[0]
        if base == "UTCHMMA":
            c["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
        elif op.startswith("MUFU.EX2"):
            c["MUFU.EX2"] += 1
        elif base in ("UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "REDG", "FFMA2", "FFMA", "HMMA"):
            c[base] += 1
print(f"# {lib}: SASS opcode counts per kernel (sm_100a); columns: " + " ".join(KEYS))
tot = collections.Counter()
for k, c in counts.items():
    name = demangle(k)
    name = re.sub(r"\(.*", "", name)[:110]
    print(f"{name:110s} " + " ".join(f"{c.get(x, 0):6d}" for x in KEYS))
    tot.update(c)
print(f"{'ALL KERNELS':110s} " + " ".join(f"{tot.get(x, 0):6d}" for x in KEYS))
