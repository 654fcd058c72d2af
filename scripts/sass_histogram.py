#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libdynare_b200.so (cuobjdump -sass; runs without a GPU).
Usage: python scripts/sass_histogram.py [out.json]      default profiles/r02_sass_opcodes.json"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dynare.jl_b200", "libdynare_b200.so")
WATCH = ("DFMA", "DMUL", "DADD", "DMMA", "MUFU", "UBLKCP", "SYNCS", "SHFL", "LDS", "STS", "LDG", "STG", "ST", "LD", "BAR",
         "ATOM", "RED", "MEMBAR", "UTMALDG", "UTCHMMA", "HMMA", "CALL")


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_opcodes.json")
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = kernels.setdefault(re.sub(r"\(.*", "", name)[:100], collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Za-z0-9_]+)*)", ln)
        if m and cur is not None:
            cur[m.group(1)] += 1
            if m.group(1) in ("ST", "LD", "MEMBAR", "DMMA", "UBLKCP", "SYNCS") and m.group(2):
                cur[m.group(1) + m.group(2)] += 1
    res = {}
    for k, c in kernels.items():
        tot = sum(v for kk, v in c.items() if "." not in kk)
        row = {"instructions": tot}
        for w in WATCH:
            if c.get(w):
                row[w] = c[w]
        for kk, v in c.items():
            if "." in kk and (".SYS" in kk or kk.startswith("DMMA") or kk.startswith("UBLKCP") or kk.startswith("SYNCS")):
                row[kk] = v
        res[k] = row
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    for k in ("dyb::cr_kernel<dyb::Model_fs2000, 4>", "dyb::kalman_kernel<dyb::Model_fs2000>", "dyb::kalman_dmma_large_kernel<true>",
              "dyb::smc_accept_publish_kernel", "dyb::smc_w1_kernel"):
        for name, row in res.items():
            if name.startswith(k):
                print(name, row)
    print(len(res), "kernels ->", out)


if __name__ == "__main__":
    main()
