"""Per-kernel SASS instruction counts of libendo_b200.so: evidence of tcgen05 (UTCHMMA/UTCQMMA), TMEM (LDTM), TMA tensor loads /
stores (UTMALDG / UTMASTG), TMA bulk copies (UBLKCP), L2 reductions (REDG) and cp.async (LDGSTS).
    python tools/sass_counts.py > profiles/r2_sass_counts.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "endoscopydepthestimation-pytorch_b200", "libendo_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "REDG", "LDGSTS", "SYNCS", "ATOMG", "RED."]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name).replace("endo::", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k in KEYS:
        if re.search(r"\b" + re.escape(k), line):
            counts[cur][k] += 1
rev = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
print(f"# SASS instruction counts per kernel, libendo_b200.so built from {rev} (`cuobjdump -sass`, sm_100a)\n")
print("UTCHMMA = tcgen05.mma kind::tf32 / f16, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG / UTMASTG = cp.async.bulk.tensor load / store,")
print("UBLKCP = cp.async.bulk (TMA bulk copy), REDG = red.global (L2 reduction), LDGSTS = cp.async, SYNCS = mbarrier operations.\n")
cols = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "REDG", "LDGSTS", "SYNCS"]
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---:|" * len(cols))
tot = collections.Counter()
for name, c in counts.items():
    if not any(c[k] for k in cols):
        continue
    print(f"| `{name[:70]}` | " + " | ".join(str(c[k]) for k in cols) + " |")
    tot.update(c)
print("| **total** | " + " | ".join(str(tot[k]) for k in cols) + " |")
