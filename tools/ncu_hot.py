#!/usr/bin/env python
"""Hot SASS instructions of one kernel of an ncu report (source page): sample share, executed count, opcode histogram.
    python tools/ncu_hot.py <report.ncu-rep> <kernel-name-substring> [top]"""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the export concatenates one table per kernel instance: take the first
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    try:
        data.append((int(r[si] or 0), r[1].strip(), int(r[ie] or 0)))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1
print(f"kernel {kern}: {len(data)} SASS instructions, {sum(d[2] for d in data)} warp-instructions executed, {tot} samples")
for s, src, e in sorted(data, key=lambda t: -t[0])[:top]:
    print(f"{100 * s / tot:5.1f}%  {e:10d}  {src[:100]}")
ops = collections.Counter()
for s, src, e in data:
    parts = src.split()
    op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
    ops[op.split(".")[0]] += e
print("executed by opcode:", ops.most_common(16))
