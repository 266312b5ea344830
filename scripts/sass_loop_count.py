#!/usr/bin/env python
"""Static size of the packed fill's chunk loop in the SASS of a built library (offline check before spending GPU time).
usage: python scripts/sass_loop_count.py smoothxg_b200/lib/libpoa_b200.so"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = txt.split("Function : ")
fn = next(f for f in funcs if f.startswith("_Z21poa_b200_block_kernelILi1E"))
ins = []
for line in fn.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_ix = {a: k for k, (a, _) in enumerate(ins)}
# the scan's widest shuffle of the single-warp packed fill (first occurrence in the function)
scan = [k for k, (_, s) in enumerate(ins) if "STS.128" in s]  # only the packed fill keeps rows in shared memory
cands = []
for k, (a, s) in enumerate(ins):
    m = re.search(r"BRA\s+(?:P\d, )?0x([0-9a-f]+)", s)
    if not m or "BRA.DIV" in s:
        continue
    t = int(m.group(1), 16)
    if t < a and t in addr_ix:
        lo, hi = addr_ix[t], k
        if any(lo <= sc <= hi for sc in scan) and any(x.startswith("SHFL.UP") for _, x in ins[lo:hi + 1]):
            cands.append((lo, hi))
# innermost loops only, in address order: global-mode instantiation first, local-mode second
inner = [c for c in cands if not any(o != c and c[0] <= o[0] and o[1] <= c[1] for o in cands)]
inner.sort()
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
best = inner[which]
lo, hi = best
body = ins[lo:hi + 1]
ops = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", s).split()[0].split(".")[0] for _, s in body)
print(f"{lib}: chunk loop {hi - lo + 1} instructions (0x{ins[lo][0]:x}..0x{ins[hi][0]:x})")
print("  " + ", ".join(f"{k} {v}" for k, v in ops.most_common(16)))
