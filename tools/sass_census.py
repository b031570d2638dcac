"""Per-kernel census of the Blackwell-specific SASS in the shipped library (what proves tcgen05 / TMEM / TMA use).

    python tools/sass_census.py [camc2v_b200/libcamc2v_b200_fp16.so] > profiles/sass_census.txt

Runs `cuobjdump -sass` (CPU box, no GPU needed) and counts, per kernel: UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st),
UTMALDG / UTMASTG (TMA tensor loads / stores), UTCBAR (tcgen05.commit), SYNCS (mbarrier), HMMA (warp-level mma.sync), MUFU.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "camc2v_b200", "libcamc2v_b200_fp16.so")
KEYS = ["UTC*MMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "LDSM", "MUFU", "FFMA2", "total"]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", o.replace("(int)", "").replace("(bool)", "").replace("void ", "")) for o in out]


txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        op = m.group(1)
        c = counts[cur]
        c["total"] += 1
        if re.match(r"UTC[A-Z]*MMA", op):
            c["UTC*MMA"] += 1
        elif op in ("LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "LDSM", "MUFU", "FFMA2"):
            c[op] += 1
names = list(counts)
pretty = demangle(names)
print(f"# {os.path.relpath(lib, ROOT)}: {len(names)} kernels; columns = static instruction counts in the sm_100a SASS")
print(f"{'kernel':88s} " + " ".join(f"{k:>8s}" for k in KEYS))
tot = collections.Counter()
for n, pn in sorted(zip(names, pretty), key=lambda t: t[1]):
    c = counts[n]
    tot.update(c)
    print(f"{pn[:88]:88s} " + " ".join(f"{c[k]:8d}" for k in KEYS))
print(f"{'ALL':88s} " + " ".join(f"{tot[k]:8d}" for k in KEYS))
