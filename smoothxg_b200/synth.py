"""Seeded synthetic POA block batches.

Each block: one uniform-random ACGT base sequence of length L; each of S copies mutates it
independently at divergence d (60 % substitution, 20 % 1-bp insertion, 20 % 1-bp deletion);
copies are ordered longest-first, the order smoothxg hands ranges to the POA engine
(reference src/blocks.cpp:204-219, src/breaks.cpp:314-327); dedup weights default to 1.
`indel_prob`/`indel_len` add one long insertion or deletion per copy, because real blocks carry
predecessor edges hundreds of rows long that 1-bp events never produce.

The layout is the flat batch the C-ABI takes (include/poa_b200.h): codes 0..3 = ACGT, 4 = N.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class PoaBatch:
    block_seq_off: np.ndarray  # int64 [n_blocks+1] -> index into seq_len/weight
    seq_len: np.ndarray        # int32 [n_seqs]
    seq_off: np.ndarray        # int64 [n_seqs+1] -> index into bases
    bases: np.ndarray          # uint8 [total]
    weight: np.ndarray         # int32 [n_seqs]

    @property
    def n_blocks(self) -> int:
        return int(self.block_seq_off.shape[0] - 1)

    @property
    def n_seqs(self) -> int:
        return int(self.seq_len.shape[0])

    def block(self, b: int):
        """(seq_len, bases, weight) views of block b."""
        s0, s1 = int(self.block_seq_off[b]), int(self.block_seq_off[b + 1])
        o0, o1 = int(self.seq_off[s0]), int(self.seq_off[s1])
        return self.seq_len[s0:s1], self.bases[o0:o1], self.weight[s0:s1]

    def block_seqs(self, b: int):
        lens, bases, _ = self.block(b)
        out, o = [], 0
        for n in lens:
            out.append(bases[o:o + int(n)])
            o += int(n)
        return out

    def select(self, idx) -> "PoaBatch":
        """Sub-batch with the given block ids, in the given order (vectorised: a shard of a 10 000-block batch in milliseconds)."""
        idx = np.asarray(idx, dtype=np.int64).reshape(-1)
        s0, s1 = self.block_seq_off[idx], self.block_seq_off[idx + 1]
        nseq = s1 - s0
        bso = np.zeros(idx.shape[0] + 1, dtype=np.int64)
        np.cumsum(nseq, out=bso[1:])
        seq_idx = np.repeat(s0 - bso[:-1], nseq) + np.arange(bso[-1], dtype=np.int64)
        seq_len = np.ascontiguousarray(self.seq_len[seq_idx], dtype=np.int32)
        seq_off = np.zeros(seq_len.shape[0] + 1, dtype=np.int64)
        np.cumsum(seq_len, out=seq_off[1:])
        b0 = self.seq_off[s0]
        nb = self.seq_off[s1] - b0  # a block's sequences are consecutive, so its bases are one contiguous run
        boff = np.zeros(idx.shape[0] + 1, dtype=np.int64)
        np.cumsum(nb, out=boff[1:])
        base_idx = np.repeat(b0 - boff[:-1], nb) + np.arange(boff[-1], dtype=np.int64)
        return PoaBatch(bso, seq_len, seq_off, np.ascontiguousarray(self.bases[base_idx], dtype=np.uint8),
                        np.ascontiguousarray(self.weight[seq_idx], dtype=np.int32))

    @staticmethod
    def from_blocks(blocks) -> "PoaBatch":
        """blocks: iterable of (list of uint8 code arrays, weights or None)."""
        bso, lens, chunks, wts = [0], [], [], []
        for seqs, w in blocks:
            for i, s in enumerate(seqs):
                s = np.asarray(s, dtype=np.uint8)
                lens.append(s.shape[0])
                chunks.append(s)
                wts.append(1 if w is None else int(w[i]))
            bso.append(len(lens))
        seq_len = np.asarray(lens, dtype=np.int32)
        seq_off = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(seq_len, out=seq_off[1:])
        bases = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint8)
        return PoaBatch(np.asarray(bso, dtype=np.int64), seq_len, seq_off,
                        np.ascontiguousarray(bases, dtype=np.uint8), np.asarray(wts, dtype=np.int32))

    @staticmethod
    def from_strings(blocks) -> "PoaBatch":
        """blocks: iterable of lists of ASCII strings; encoded like ab_char26_table
        (reference deps/abPOA/src/abpoa_seq.c:15-32): ACGT/acgt -> 0..3, U/u -> 3, else 4."""
        return PoaBatch.from_blocks([([encode(s) for s in seqs], None) for seqs in blocks])


_ENC = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _ENC[ord(_c)] = _i
    _ENC[ord(_c.lower())] = _i
_ENC[ord("U")] = 3
_ENC[ord("u")] = 3
_ENC[0:4] = np.arange(4, dtype=np.uint8)


def encode(s: str) -> np.ndarray:
    return _ENC[np.frombuffer(s.encode("ascii"), dtype=np.uint8)]


def decode(codes) -> str:
    return "".join("ACGTN"[int(c)] for c in codes)


def _mutate(rng, base: np.ndarray, d: float, indel_prob: float, indel_len) -> np.ndarray:
    L = base.shape[0]
    u = rng.random(L)
    kind = rng.random(L)
    mut = u < d
    sub = mut & (kind < 0.6)
    ins = mut & (kind >= 0.6) & (kind < 0.8)
    dele = mut & (kind >= 0.8)
    b = base.copy()
    b[sub] = (b[sub] + 1 + rng.integers(0, 3, int(sub.sum()), dtype=np.uint8)) % 4
    reps = np.ones(L, dtype=np.int64)
    reps[ins] = 2
    reps[dele] = 0
    out = np.repeat(b, reps)
    # the inserted base is the second copy of each doubled position: overwrite it with a random base
    if ins.any():
        ends = np.cumsum(reps)
        pos = ends[ins] - 1
        out[pos] = rng.integers(0, 4, pos.shape[0], dtype=np.uint8)
    if indel_prob > 0 and rng.random() < indel_prob and out.shape[0] > 4:
        lo, hi = indel_len
        n = int(rng.integers(lo, hi + 1))
        at = int(rng.integers(1, out.shape[0] - 1))
        if rng.random() < 0.5:
            out = np.concatenate([out[:at], rng.integers(0, 4, n, dtype=np.uint8), out[at:]])
        else:
            n = min(n, out.shape[0] - at - 1)
            out = np.concatenate([out[:at], out[at + n:]])
    return out.astype(np.uint8)


def make_batch(n_blocks: int, n_seqs: int, length: int, divergence: float = 0.02, seed: int = 1,
               indel_prob: float = 0.0, indel_len=(50, 500), n_frac: float = 0.0,
               dup_weights: bool = False) -> PoaBatch:
    """Deterministic synthetic batch (numpy PCG64 seeded with `seed`).

    n_frac > 0 replaces that fraction of bases with N (code 4) to exercise the zero-score row/column
    (reference deps/abPOA/src/abpoa_align.c:19-22).  dup_weights draws dedup multiplicities 1..3.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    blocks = []
    for _ in range(n_blocks):
        base = rng.integers(0, 4, length, dtype=np.uint8)
        seqs = [_mutate(rng, base, divergence, indel_prob, indel_len) for _ in range(n_seqs)]
        if n_frac > 0:
            for s in seqs:
                s[rng.random(s.shape[0]) < n_frac] = 4
        order = sorted(range(n_seqs), key=lambda i: (-seqs[i].shape[0], i))
        seqs = [seqs[i] for i in order]
        w = rng.integers(1, 4, n_seqs) if dup_weights else None
        blocks.append((seqs, w))
    return PoaBatch.from_blocks(blocks)


# BASELINE.json configs 2-4
CONFIGS = {
    "config1_1000x16x1k": dict(n_blocks=1000, n_seqs=16, length=1000),
    "config2_10000x32x2k": dict(n_blocks=10000, n_seqs=32, length=2000),
    "config3_deep_100x256x8k": dict(n_blocks=100, n_seqs=256, length=8000),
}
