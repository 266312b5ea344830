#!/usr/bin/env python
"""Small POA batches for compute-sanitizer runs (memcheck / racecheck): one engine per warps-per-block setting, global banded
(packed 16-bit fill), local (exact BFS order) and an int32 / affine case (generic fill), each checked against the oracle.
usage: compute-sanitizer --tool racecheck python scripts/sanitize_case.py <warps>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from oracle.oracle import Oracle, make_params as oparams  # noqa: E402
from smoothxg_b200 import engine, synth  # noqa: E402
from tests.helpers import view_to_dump  # noqa: E402

warps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
eng = engine.PoaEngine(device=0, emit_cigar=True, warps_per_block=warps)
ora = Oracle()
cases = [(dict(n_blocks=3, n_seqs=6, length=700, seed=5, indel_prob=0.4, indel_len=(20, 150)), dict(out_msa=True)),
         (dict(n_blocks=2, n_seqs=5, length=300, seed=6), dict(local=True)),
         (dict(n_blocks=2, n_seqs=5, length=300, seed=7), dict(gap_open2=0, gap_ext2=0))]
for kw, pk in cases:
    batch = synth.make_batch(**kw)
    res = eng.run_batch(batch, engine.make_params(**pk))
    for b in range(batch.n_blocks):
        want = ora.poa_block(oparams(**pk), *batch.block(b))
        assert np.array_equal(want.compare_part(), view_to_dump(res.block(b)).compare_part()), (kw, b)
    print("ok", warps, pk, res.stats()["kernel_ms"], "ms")
    res.close()
eng.close()
