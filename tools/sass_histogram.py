#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass): which kernels are tcgen05 / TMEM / TMA
code and which still run on legacy mma.sync.   python tools/sass_histogram.py [lib.so] > profiles/rNN_sass_histogram.txt

  UTCHMMA   tcgen05.mma (kind::f16), `.2CTA` = cta_group::2        LDTM / STTM   tcgen05.ld / st (TMEM)
  UTMALDG   TMA tensor load     UTMASTG  TMA tensor store          UTCBAR        tcgen05.commit -> mbarrier
  UBLKCP    cp.async.bulk       SYNCS    mbarrier ops              HMMA          legacy mma.sync
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "self-guided-diffusion-models_b200", "libsgdm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UBLKCP", "SYNCS", "HMMA", "MUFU",
        "LDGSTS", "LDSM", "FFMA", "total"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["total"] += 1
        base = op.split(".")[0]
        per[cur][base] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            per[cur]["UTCHMMA.2CTA"] += 1


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception:
        return n


print(f"# {os.path.basename(lib)}: SASS opcode counts per kernel (sm_100a); columns: " + " ".join(KEYS))
w = max(len(demangle(k)) for k in per) if per else 10
print(f"{'kernel':{min(w, 70)}s} " + " ".join(f"{k:>8s}" if len(k) <= 8 else f"{k:>12s}" for k in KEYS))
tot = collections.Counter()
for k, c in per.items():
    name = demangle(k)[:70]
    print(f"{name:{min(w, 70)}s} " + " ".join(f"{c[x]:8d}" if len(x) <= 8 else f"{c[x]:12d}" for x in KEYS))
    tot.update(c)
print(f"{'ALL':{min(w, 70)}s} " + " ".join(f"{tot[x]:8d}" if len(x) <= 8 else f"{tot[x]:12d}" for x in KEYS))
