#!/usr/bin/env python
"""Generate tests/golden/deep_golden.npz: full-size single blocks of the BASELINE.json shapes that the small golden set
only covers in miniature, run through the UNMODIFIED vendored abPOA (oracle/_ref):
  * deep_256x8kb  -- one configs[3] block (256 sequences x 8 kb, 2 % divergence, global, adaptive band): the graph passes
                     16 361 rows part-way, so abPOA switches from int16 to int32 scores mid-block (abpoa_align_simd.c:1293-1302);
  * local_32x2kb  -- one configs[2]-shaped block in LOCAL mode (what plain -A gives, src/main.cpp:487: unbanded, BFS row order).
Inputs are regenerated from the seeds (smoothxg_b200/synth.py), so only the dumps are stored (compressed).
  python tests/golden/make_deep_golden.py      (about a minute of CPU)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import RefAbpoa, make_params  # noqa: E402
from smoothxg_b200.synth import make_batch  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "deep_golden.npz")
CASES = {
    "deep_256x8kb": (dict(n_blocks=1, n_seqs=256, length=8000, divergence=0.02, seed=3001), dict()),
    "local_32x2kb": (dict(n_blocks=1, n_seqs=32, length=2000, divergence=0.02, seed=3002), dict(local=True)),
}

if __name__ == "__main__":
    ref = RefAbpoa()
    out = {}
    for name, (kw, pk) in CASES.items():
        batch = make_batch(**kw)
        d = ref.poa_block(make_params(**pk), *batch.block(0), instrument=True)
        out[name] = d.raw
        print(name, "nodes", d.n_node, "in-band cells", d.inband_cells, "dump words", d.raw.shape[0], flush=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
