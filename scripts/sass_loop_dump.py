#!/usr/bin/env python
"""Print the SASS of the packed fill's chunk loop (see sass_loop_count.py). usage: sass_loop_dump.py lib.so [which]"""
import re, subprocess, sys
lib = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn = next(f for f in txt.split("Function : ") if f.startswith("_Z21poa_b200_block_kernelILi1E"))
ins = [(int(m.group(1), 16), m.group(2).strip()) for m in (re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l) for l in fn.splitlines()) if m]
ix = {a: k for k, (a, _) in enumerate(ins)}
scan = [k for k, (_, s) in enumerate(ins) if "STS.128" in s]
cands = []
for k, (a, s) in enumerate(ins):
    m = re.search(r"BRA\s+(?:P\d, )?0x([0-9a-f]+)", s)
    if m and "BRA.DIV" not in s:
        t = int(m.group(1), 16)
        if t < a and t in ix:
            lo, hi = ix[t], k
            if any(lo <= c <= hi for c in scan) and any(x.startswith("SHFL.UP") for _, x in ins[lo:hi + 1]):
                cands.append((lo, hi))
inner = sorted(c for c in cands if not any(o != c and c[0] <= o[0] and o[1] <= c[1] for o in cands))
lo, hi = inner[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
for a, s in ins[lo:hi + 1]:
    print(f"{a:06x}  {s}")
