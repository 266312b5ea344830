#!/usr/bin/env python
"""Generate tests/golden/mash_golden.npz from the UNMODIFIED mkmh/rkmh headers of the reference
(oracle/_ref/libmash_ref.so = oracle/ref_mash_shim.cpp + deps/mkmh/{mkmh,rkmh}.hpp + murmur3.cpp compiled where they lie).
Run in the build container (needs /root/reference); the npz is committed so the GPU box can pin the oracle and the CUDA
path to the reference.  Upstream holds no known-answer vectors for this path (deps/mkmh/mkmh_test.cpp exercises other
entry points), so these are ours: per case the kept count, est_identity_threshold, every pair's estimated identity
and the sorted hash list of the first string."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.mash import MashRef  # noqa: E402
from tests.mash_cases import make_cases  # noqa: E402

ref = MashRef()
out = {}
names = []
for name, k, seqs in make_cases():
    kept, thr, ident, _ = ref.block(seqs, k)
    names.append(name)
    out[f"{name}/kmer"] = np.int32(k)
    out[f"{name}/kept"] = np.int32(kept)
    out[f"{name}/threshold"] = np.float32(-1.0 if thr is None else thr)
    out[f"{name}/pair_identity"] = ident
    out[f"{name}/hashes0"] = ref.hashes(seqs[0], k) if seqs and len(seqs[0]) > k else np.zeros(0, dtype=np.uint64)
    out[f"{name}/n_seq"] = np.int32(len(seqs))
    out[f"{name}/seqs"] = np.frombuffer("\n".join(seqs).encode(), dtype=np.uint8)  # newline-joined
out["names"] = np.frombuffer("\n".join(names).encode(), dtype=np.uint8)
path = os.path.join(ROOT, "tests", "golden", "mash_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes,", len(names), "cases")
