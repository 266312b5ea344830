#!/usr/bin/env python
"""CPU fuzz of the identity estimate: seeded block sets (tests/mash_cases.py, every branch of src/smooth.cpp:1982-2023 and
rkmh::compare) through the oracle restatement, the unmodified mkmh/rkmh headers (oracle/_ref) and the host replay of the
device logic (tests/emu/emu_mash.cpp).  usage: python scripts/fuzz_mash.py [n_seeds] [first_seed]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from oracle.mash import MashOracle, MashRef, ref_available  # noqa: E402
from tests.mash_cases import make_cases  # noqa: E402
from tests.test_mash import _emu_lib  # noqa: E402

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 12
first = int(sys.argv[2]) if len(sys.argv) > 2 else 100
o, lib = MashOracle(), _emu_lib()
r = MashRef() if ref_available() else None
n = 0
for seed in range(first, first + n_seeds):
    for name, k, seqs in make_cases(seed=seed, n=24):
        ko, to, io, co = o.block(seqs, k)
        if r is not None:
            kr, tr, ir, _ = r.block(seqs, k)
            assert ko == kr and to == tr and np.array_equal(io, ir), (seed, name)
        keep = [s for s in seqs if len(s) >= 8 * k]
        if len(keep) >= 2:
            bufs = [s.encode() for s in keep]
            arr = (C.c_char_p * len(bufs))(*bufs)
            lens = (C.c_int * len(bufs))(*[len(b) for b in bufs])
            got = np.zeros(len(co), dtype=np.uint32)
            lib.emu_mash_block_common(len(bufs), arr, lens, k, got.ctypes.data_as(C.c_void_p))
            assert np.array_equal(got.astype(np.uint64), co), (seed, name)
        n += 1
print(f"mash fuzz ok: {n} blocks ({'with' if r is not None else 'without'} the unmodified reference)")
