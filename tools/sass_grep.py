"""Per-kernel counts of the Blackwell-native SASS mnemonics in the built library (what proves tcgen05 / TMEM / TMA / cluster use):
    python tools/sass_grep.py > profiles/r02_sass_grep.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "dreamer_v3_b200", "libdv3.so")
MN = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UCGABAR", "HMMA", "REDG")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("dv3::", "").replace("void ", "")
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k in MN:
        if re.search(r"\b" + re.escape(k), line):
            counts[cur][k.rstrip(".")] += 1
cols = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UCGABAR", "HMMA", "REDG"]
print("# cuobjdump -sass dreamer_v3_b200/libdv3.so: occurrences per kernel (tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, TMA = UTMALDG /")
print("# UTMASTG / UTMAREDG, bulk copy (shared::cta -> shared::cluster exchange) = UBLKCP, cluster barrier = UCGABAR, mma.sync = HMMA,")
print("# red.global = REDG); kernels without any of them omitted")
print(f"{'kernel':90s} " + " ".join(f"{c:>8s}" for c in cols))
for k, c in counts.items():
    if not any(c[x] for x in cols[:8]):
        continue
    print(f"{k[:90]:90s} " + " ".join(f"{c[x]:8d}" for x in cols))
print(f"# {len(counts)} kernels in the library; {sum(1 for c in counts.values() if c['UTCHMMA'])} issue tcgen05.mma; {sum(1 for c in counts.values() if c['HMMA'])} use mma.sync (HMMA)")
