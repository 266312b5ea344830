#!/usr/bin/env python
"""GPU fuzz of the CUDA path through the C ABI: random scoring parameters (convex / affine / linear, presets, odd extension orders),
modes (global banded / unbanded / local), warps per block (auto / 1 / 2 / 4 / 8) and random blocks (indels, N runs, dedup weights,
a few long-graph blocks beyond abPOA's int16 rule), many blocks per launch, every block compared with the oracle restatement
(graph, edge order, weights, paths, consensus, MSA, scores, cigars, in-band cells) and a quarter of them with the unmodified abPOA.
usage: python scripts/fuzz_gpu.py [n_param_sets] [seed]     (about 8 blocks per parameter set)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from oracle.oracle import Oracle  # noqa: E402
from scripts.fuzz_emu import random_params  # noqa: E402
from smoothxg_b200 import engine, synth  # noqa: E402
from tests.helpers import first_diff, view_to_dump  # noqa: E402


def main():
    n_sets = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    ora = Oracle()
    ref = None
    try:
        from oracle.oracle import RefAbpoa
        ref = RefAbpoa()
    except Exception:  # noqa: BLE001
        pass
    engines = {w: engine.PoaEngine(device=0, emit_cigar=True, warps_per_block=w) for w in (0, 1, 2, 4, 8)}
    t0, n_blocks, n_ref, n_unsup, n_long = time.time(), 0, 0, 0, 0
    for s in range(n_sets):
        p = random_params(rng)
        ep = engine.PoaParams(p.match, p.mismatch, p.gap_open1, p.gap_ext1, p.gap_open2, p.gap_ext2, p.align_mode, p.wb, p.wf, p.out_cons, p.out_msa)
        blocks = []
        for _ in range(int(rng.integers(4, 12))):
            L = int(rng.integers(20, 1800))
            b = synth.make_batch(n_blocks=1, n_seqs=int(rng.integers(2, 12)), length=L, divergence=float(rng.choice([0.0, 0.02, 0.1, 0.3])),
                                 seed=int(rng.integers(1 << 30)), indel_prob=float(rng.choice([0.0, 0.3, 0.8])), indel_len=(5, max(6, L // 3)),
                                 n_frac=float(rng.choice([0.0, 0.0, 0.05])), dup_weights=bool(rng.integers(0, 2)))
            blocks.append((b.block_seqs(0), b.block(0)[2]))
        if s % 10 == 3:  # a long-graph block: unrelated 4.5 kb sequences, the graph passes abPOA's int16 row limit (16 361)
            b = synth.make_batch(n_blocks=1, n_seqs=12, length=4500, divergence=0.75, seed=int(rng.integers(1 << 30)))
            blocks.append((b.block_seqs(0), b.block(0)[2])); n_long += 1
        batch = synth.PoaBatch.from_blocks(blocks)
        w = int(rng.choice([0, 1, 2, 4, 8]))
        try:
            res = engines[w].run_batch(batch, ep)
        except engine.PoaError as e:
            if e.code == engine.EUNSUP:
                n_unsup += 1
                continue
            raise
        for b in range(batch.n_blocks):
            want = ora.poa_block(p, *batch.block(b))
            got = view_to_dump(res.block(b))
            tag = f"set {s} seed {seed} warps {w} block {b}: params {[getattr(p, f) for f, _ in p._fields_]}"
            assert np.array_equal(got.compare_part(), want.compare_part()), f"{tag}: {first_diff(want, got)}"
            if ref is not None and (n_blocks + b) % 4 == 0:
                r = ref.poa_block(p, *batch.block(b))
                assert r is not None and np.array_equal(r.compare_part(), want.compare_part()), f"oracle vs abPOA: {tag}"
                n_ref += 1
        n_blocks += batch.n_blocks
        res.close()
    print(f"gpu fuzz ok: {n_sets} parameter sets, {n_blocks} blocks ({n_long} long-graph blocks, {n_unsup} rejected parameter sets, "
          f"{n_ref} blocks also checked against unmodified abPOA) in {time.time() - t0:.1f} s")


if __name__ == "__main__":
    main()
