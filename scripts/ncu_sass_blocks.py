#!/usr/bin/env python
"""Where a kernel's instructions and stall samples go, from the SASS view of an ncu report:
opcode mix and contiguous SASS blocks by execution count.  Usage: python scripts/ncu_sass_blocks.py rep.ncu-rep [units]
(units = the number of work units, e.g. draw-periods, to normalise the counts by)."""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[1]
si, ii, wi = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
data = [(int(r[ii]), int(r[wi]), r[si].strip(), [int(r[i] or 0) for i in stall_cols]) for r in rows[2:] if len(r) > wi and r[ii].isdigit()]
tot = sum(d[0] for d in data); tw = sum(d[1] for d in data)
print(len(data), "sass instrs; executed", tot, f"({tot / units:.0f} per unit); samples", tw)
ops = collections.Counter()
for n, w, s, _ in data:
    ops[re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", s).group(1)] += n
print("  ".join(f"{k} {v / tot * 100:.1f}% ({v / units:.0f})" for k, v in ops.most_common(16)))
st = collections.Counter()
for d in data:
    for i, c in zip(stall_cols, d[3]):
        st[h[i]] += c
print("  ".join(f"{k[6:]} {v / tw * 100:.1f}%" for k, v in st.most_common(8)))
blk = []
def flush():
    if not blk:
        return
    n = sum(b[0] for b in blk); w = sum(b[1] for b in blk)
    if n / tot > 0.01 or w / tw > 0.01:
        c = collections.Counter(re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", b[2]).group(1) for b in blk)
        s2 = collections.Counter()
        for b in blk:
            for i, cc in zip(stall_cols, b[4]):
                s2[h[i][6:]] += cc
        print(f"idx {blk[0][3]:5d}-{blk[-1][3]:5d} len {len(blk):4d} exec/inst {blk[0][0] / units:8.2f} inst {n / tot * 100:5.1f}% smp {w / tw * 100:5.1f}%  "
              f"{dict(c.most_common(5))}  stalls {dict(s2.most_common(3))}")
prev = None
for i, (n, w, s, sc) in enumerate(data):
    if prev is not None and (n > prev * 1.3 + 10 or n < prev / 1.3 - 10):
        flush(); blk = []
    blk.append((n, w, s, i, sc)); prev = n
flush()
