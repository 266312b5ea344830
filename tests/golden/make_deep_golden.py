#!/usr/bin/env python
"""Generate tests/golden/deep_golden.npz: full-size single blocks of the BASELINE.json shapes that the small golden set
only covers in miniature, run through the UNMODIFIED vendored abPOA (oracle/_ref):
  * deep_256x8kb  -- one configs[3] block (256 sequences x 8 kb, 2 % divergence, global, adaptive band): the graph passes
                     16 361 rows part-way, so abPOA switches from int16 to int32 scores mid-block (abpoa_align_simd.c:1293-1302);
  * local_32x2kb  -- one configs[2]-shaped block in LOCAL mode (what plain -A gives, src/main.cpp:487: unbanded, BFS row order).
Inputs are regenerated from the seeds (smoothxg_b200/synth.py).  A deep block's dump is 26 MB, so the file stores, per
case, the dump header, the small sections verbatim (per-sequence scores, cigar and path lengths, consensus) and a SHA-256 of
every section (graph, edge lists, weights, aligned groups, paths, cigars): equality of all digests is bit-exact parity, and
a mismatch names the section.
  python tests/golden/make_deep_golden.py      (about a minute of CPU)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import RefAbpoa, make_params  # noqa: E402
from smoothxg_b200.synth import make_batch  # noqa: E402
from tests.golden_io import DEEP_CASES as CASES, digest_dump  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "deep_golden.npz")
if __name__ == "__main__":
    ref = RefAbpoa()
    out = {}
    for name, (kw, pk) in CASES.items():
        batch = make_batch(**kw)
        d = ref.poa_block(make_params(**pk), *batch.block(0), instrument=True)
        out.update(digest_dump(name, d))
        print(name, "nodes", d.n_node, "in-band cells", d.inband_cells, "dump words", d.raw.shape[0], flush=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
