#!/usr/bin/env python
"""CPU fuzz of the device logic: random scoring parameters (convex / affine / linear, presets, odd extension orders), modes
(global banded / unbanded / local) and small random blocks (indels, N runs, duplicates) through the emulated device code
(tests/emu: 1 lane and 32 lock-step lanes) against the oracle restatement, and the oracle against the unmodified abPOA where
oracle/_ref exists.  usage: python scripts/fuzz_emu.py [n_cases] [seed]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from oracle.oracle import Oracle, PdParams, _Checker  # noqa: E402
from smoothxg_b200 import synth  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def random_params(rng) -> PdParams:
    kind = rng.integers(0, 10)
    if kind < 4:      # convex, smoothxg-like
        m, n = 1, int(rng.integers(2, 20))
        o1, e1 = int(rng.integers(2, 40)), int(rng.integers(1, 4))
        o2, e2 = int(rng.integers(o1 + 1, 90)), int(rng.integers(1, 3))
    elif kind < 6:    # convex, odd: second piece steeper / equal extensions / larger match
        m, n = int(rng.integers(1, 4)), int(rng.integers(1, 9))
        o1, e1 = int(rng.integers(1, 12)), int(rng.integers(1, 5))
        o2, e2 = int(rng.integers(1, 30)), int(rng.integers(1, 5))
    elif kind < 8:    # affine (gap_open2 == 0)
        m, n = int(rng.integers(1, 3)), int(rng.integers(1, 9))
        o1, e1, o2, e2 = int(rng.integers(1, 20)), int(rng.integers(1, 4)), 0, int(rng.integers(0, 3))
    else:             # linear (gap_open1 == 0)
        m, n = int(rng.integers(1, 3)), int(rng.integers(1, 9))
        o1, e1, o2, e2 = 0, int(rng.integers(1, 6)), int(rng.integers(0, 10)), int(rng.integers(0, 3))
    mode = int(rng.integers(0, 3))  # 0 global banded, 1 global unbanded, 2 local
    return PdParams(m, n, o1, e1, o2, e2, 1 if mode == 2 else 0, 311 if mode == 0 else -1, 0.03, 1, int(rng.integers(0, 2)))


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    ora = Oracle()
    emus = {"1 lane": _Checker(C.CDLL(os.path.join(ROOT, "tests/emu/_build/libpoa_emu.so")), "emu_poa_block", "emu_free"),
            "32 lanes": _Checker(C.CDLL(os.path.join(ROOT, "tests/emu/_build/libpoa_emu32.so")), "emu_poa_block", "emu_free")}
    if os.environ.get("FUZZ_X4"):  # four warps of 32 lanes per POA block: the multi-warp fills, cross-warp barriers included
        emus["4 x 32 lanes"] = _Checker(C.CDLL(os.path.join(ROOT, "tests/emu/_build/libpoa_emu32x4.so")), "emu_poa_block", "emu_free")
    ref = None
    try:
        from oracle.oracle import RefAbpoa
        ref = RefAbpoa()
    except Exception:
        pass
    t0, n_ref, n_unsup = time.time(), 0, 0
    for c in range(n_cases):
        p = random_params(rng)
        L = int(rng.integers(20, int(os.environ.get("FUZZ_LMAX", "700"))))  # FUZZ_LMAX=2500: rows of several chunks with a real band
        batch = synth.make_batch(n_blocks=1, n_seqs=int(rng.integers(2, 9)), length=L, divergence=float(rng.choice([0.0, 0.02, 0.1, 0.3])),
                                 seed=int(rng.integers(1 << 30)), indel_prob=float(rng.choice([0.0, 0.3, 0.8])), indel_len=(5, max(6, L // 3)),
                                 n_frac=float(rng.choice([0.0, 0.0, 0.05])), dup_weights=bool(rng.integers(0, 2)))
        blk = batch.block(0)
        want = ora.poa_block(p, *blk)
        tag = f"case {c} seed {seed}: params {[getattr(p, f) for f, _ in p._fields_]} L={L}"
        if want is None:
            n_unsup += 1
            continue
        for name, emu in emus.items():
            got = emu.poa_block(p, *blk)
            if got is None:   # parameters the engine rejects (EUNSUP) are not a parity failure
                n_unsup += 1
                continue
            assert np.array_equal(got.compare_part(), want.compare_part()), f"{name}: {tag}"
        if ref is not None and c % 4 == 0:
            r = ref.poa_block(p, *blk)
            assert r is not None and np.array_equal(r.compare_part(), want.compare_part()), f"oracle vs abPOA: {tag}"
            n_ref += 1
    print(f"fuzz ok: {n_cases} cases ({n_unsup} rejected runs, {n_ref} checked against unmodified abPOA) in {time.time() - t0:.1f} s")


if __name__ == "__main__":
    main()
