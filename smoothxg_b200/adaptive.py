"""ctypes binding of include/mash_b200.h: the per-block identity estimate behind smoothxg's --adaptive-poa-params
(reference src/smooth.cpp:1982-2062).  Test / bench harness only -- the product is the C ABI; everything heavy runs in
the CUDA kernels of csrc/mash_b200.cu and the call fails without a GPU (no CPU fallback)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import engine


class MashStats(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("hash_ms", C.c_double), ("sort_ms", C.c_double), ("compare_ms", C.c_double),
                ("d2h_ms", C.c_double), ("host_ms", C.c_double),
                ("n_seqs_kept", C.c_int64), ("n_hashes", C.c_int64), ("n_pairs", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("kernel_launches", C.c_int32), ("n_chunks", C.c_int32)]


_bound = False


def _lib() -> C.CDLL:
    global _bound
    lib = engine.load_library()
    if not _bound:
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        lib.mash_b200_last_error.restype = C.c_char_p
        lib.mash_b200_pair_offsets.restype = i64
        lib.mash_b200_pair_offsets.argtypes = [i32, i64, vp, vp, vp]
        lib.mash_b200_block_identity.restype = C.c_int
        lib.mash_b200_block_identity.argtypes = [C.c_int, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(MashStats)]
        lib.mash_b200_preset.restype = C.c_int
        lib.mash_b200_preset.argtypes = [C.c_float, C.POINTER(i32 * 6)]
        _bound = True
    return lib


@dataclass
class FlatBlocks:
    """Blocks of ASCII strings in the flat layout the ABI takes."""
    block_seq_off: np.ndarray  # int64 [n_blocks + 1]
    seq_len: np.ndarray        # int32 [n_seqs]
    seq_off: np.ndarray        # int64 [n_seqs + 1]
    bases: np.ndarray          # uint8, ASCII

    @property
    def n_blocks(self) -> int:
        return len(self.block_seq_off) - 1

    def strings(self, b: int):
        out = []
        for s in range(int(self.block_seq_off[b]), int(self.block_seq_off[b + 1])):
            o = int(self.seq_off[s])
            out.append(self.bases[o:o + int(self.seq_len[s])].tobytes())
        return out


def flatten(blocks) -> FlatBlocks:
    bso, sl, chunks = [0], [], []
    for blk in blocks:
        for s in blk:
            b = s.encode() if isinstance(s, str) else bytes(s)
            sl.append(len(b)); chunks.append(np.frombuffer(b, dtype=np.uint8))
        bso.append(len(sl))
    seq_len = np.asarray(sl, dtype=np.int32)
    seq_off = np.zeros(len(sl) + 1, dtype=np.int64)
    np.cumsum(seq_len, out=seq_off[1:])
    bases = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint8)
    return FlatBlocks(np.asarray(bso, dtype=np.int64), seq_len, seq_off, np.ascontiguousarray(bases))


def from_codes(batch) -> FlatBlocks:
    """A synth.Batch (codes 0..4) as ASCII blocks (the strings XG would hand out)."""
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    return FlatBlocks(np.ascontiguousarray(batch.block_seq_off, dtype=np.int64), np.ascontiguousarray(batch.seq_len, dtype=np.int32),
                      np.ascontiguousarray(batch.seq_off, dtype=np.int64), np.ascontiguousarray(lut[batch.bases]))


def pair_offsets(fb: FlatBlocks, kmer: int = 17) -> np.ndarray:
    out = np.zeros(fb.n_blocks + 1, dtype=np.int64)
    _lib().mash_b200_pair_offsets(kmer, fb.n_blocks, fb.block_seq_off.ctypes.data, fb.seq_len.ctypes.data, out.ctypes.data)
    return out


def preset(threshold: float):
    """(m, n, g, e, q, c) for an estimated identity, or None: keep the user's scores (src/smooth.cpp:2026-2062)."""
    s = (C.c_int32 * 6)()
    return tuple(s) if _lib().mash_b200_preset(C.c_float(threshold), C.byref(s)) else None


def block_identity(fb: FlatBlocks, kmer: int = 17, device: int = 0, want_pairs: bool = False) -> dict:
    """est_identity_threshold per block (-1: fewer than two strings of >= 8*kmer bases), on the GPU."""
    lib = _lib()
    nb = fb.n_blocks
    thr = np.full(nb, -2.0, dtype=np.float32)
    kept = np.zeros(nb, dtype=np.int32)
    poff = pair_offsets(fb, kmer)
    common = np.zeros(max(int(poff[-1]), 1), dtype=np.uint32) if want_pairs else None
    ident = np.zeros(max(int(poff[-1]), 1), dtype=np.float32) if want_pairs else None
    st = MashStats()
    rc = lib.mash_b200_block_identity(device, kmer, nb, fb.block_seq_off.ctypes.data, fb.seq_len.ctypes.data, fb.seq_off.ctypes.data,
                                      (fb.bases if fb.bases.size else np.zeros(1, dtype=np.uint8)).ctypes.data, thr.ctypes.data, kept.ctypes.data,
                                      common.ctypes.data if want_pairs else None, ident.ctypes.data if want_pairs else None, C.byref(st))
    if rc != 0:
        raise RuntimeError(f"mash_b200_block_identity failed ({rc}): {lib.mash_b200_last_error().decode()}")
    out = {"threshold": thr, "n_kept": kept, "pair_off": poff, "stats": {k: getattr(st, k) for k, _ in MashStats._fields_}}
    if want_pairs:
        out["pair_common"] = common[:int(poff[-1])]
        out["pair_identity"] = ident[:int(poff[-1])]
    return out
