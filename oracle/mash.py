"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU checkers of the --adaptive-poa-params identity estimate.

  * `MashOracle` -> oracle/libmash_oracle.so  (our scalar C restatement, mash_oracle.c)
  * `MashRef`    -> oracle/_ref/libmash_ref.so (unmodified mkmh/rkmh headers of the reference + ref_mash_shim.cpp)

Only tests/, __graft_entry__.smoke() and bench-side CPU baselines may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _marshal(seqs):
    n = len(seqs)
    bufs = [bytes(s) if not isinstance(s, str) else s.encode() for s in seqs]
    arr = (C.c_char_p * max(n, 1))(*bufs)
    lens = (C.c_int * max(n, 1))(*[len(b) for b in bufs])
    return n, arr, lens, bufs


class _Mash:
    def __init__(self, path, prefix, has_common):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self._block = getattr(self.lib, prefix + "_block")
        self._hashes = getattr(self.lib, prefix + "_hashes")
        self._block.restype = C.c_int
        self._hashes.restype = C.c_int
        self.has_common = has_common

    def hashes(self, seq, kmer: int) -> np.ndarray:
        b = seq.encode() if isinstance(seq, str) else bytes(seq)
        out = np.zeros(max(len(b), 1), dtype=np.uint64)
        n = self._hashes(C.c_char_p(b), C.c_int(len(b)), C.c_int(kmer), out.ctypes.data_as(C.c_void_p))
        return out[:n].copy()

    def block(self, seqs, kmer: int):
        """-> (kept, threshold or None, pair identities float32[kept*(kept-1)/2], pair commons uint64 or None)"""
        n, arr, lens, _keep = _marshal(seqs)
        np_max = max(n * (n - 1) // 2, 1)
        ident = np.zeros(np_max, dtype=np.float32)
        common = np.zeros(np_max, dtype=np.uint64)
        thr = C.c_float(-1.0)
        args = [C.c_int(n), arr, lens, C.c_int(kmer), C.byref(thr), ident.ctypes.data_as(C.c_void_p)]
        if self.has_common:
            args.append(common.ctypes.data_as(C.c_void_p))
        kept = self._block(*args)
        npairs = kept * (kept - 1) // 2 if kept > 1 else 0
        return kept, (np.float32(thr.value) if kept > 1 else None), ident[:npairs].copy(), (common[:npairs].copy() if self.has_common else None)


class MashOracle(_Mash):
    def __init__(self):
        super().__init__(os.path.join(HERE, "libmash_oracle.so"), "mash", True)
        self.lib.mash_preset.restype = C.c_int

    def preset(self, threshold: float):
        s = (C.c_int * 6)()
        ok = self.lib.mash_preset(C.c_float(threshold), s)
        return tuple(s) if ok else None


class MashRef(_Mash):
    def __init__(self):
        super().__init__(os.path.join(HERE, "_ref", "libmash_ref.so"), "mash_ref", False)


def ref_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libmash_ref.so"))
