#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, legacy mma.sync -> HMMA), from `cuobjdump -sass` of the
shipped library.  Runs without a GPU.    python tools/sass_evidence.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "ivid_b200", "libivid_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "UTCBAR", "HMMA", "SYNCS", "ATOMG", "RED"]
cur = None
tab = collections.OrderedDict()
arch = set()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        cur = tab.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        for p in pats:
            if op == p or (p in ("RED", "ATOMG") and op.startswith(p)):
                cur[p] += 1
print(f"# cuobjdump -sass ivid_b200/libivid_b200.so   (arch: {', '.join(sorted(arch))})")
print("# kernel, " + ", ".join(pats))
tot = collections.Counter()
for k, c in tab.items():
    tot.update(c)
    print(f"{k}, " + ", ".join(str(c.get(p, 0)) for p in pats))
print("TOTAL, " + ", ".join(str(tot.get(p, 0)) for p in pats))
