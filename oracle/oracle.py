"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the two CPU checkers.

  * `Oracle`   -> oracle/libpoa_oracle.so   (our scalar C restatement, poa_oracle.c)
  * `RefAbpoa` -> oracle/_ref/libabpoa_ref_<isa>.so (unmodified vendored abPOA + ref_shim.c)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module;
the product (smoothxg_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# header indices, keep in sync with poa_dump.h
PD_MAGIC, PD_N_NODE, PD_N_SEQ, PD_CONS_LEN, PD_MSA_LEN, PD_MSA_ROWS, PD_N_IN_TOT, PD_N_OUT_TOT, \
    PD_N_ALN_TOT, PD_PATH_TOT, PD_CIGAR_TOT, PD_INBAND_LO, PD_INBAND_HI, PD_FULL_LO, PD_FULL_HI, \
    PD_EDGE_ROWS_LO, PD_EDGE_ROWS_HI = range(17)
PD_HEADER_LEN = 24
POA_DUMP_MAGIC = 0x504F4131


class PdParams(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open1", C.c_int32),
                ("gap_ext1", C.c_int32), ("gap_open2", C.c_int32), ("gap_ext2", C.c_int32),
                ("align_mode", C.c_int32), ("wb", C.c_int32), ("wf", C.c_float),
                ("out_cons", C.c_int32), ("out_msa", C.c_int32)]


def make_params(match=1, mismatch=4, gap_open1=6, gap_ext1=2, gap_open2=26, gap_ext2=1,
                local=False, banded=True, out_cons=True, out_msa=False) -> PdParams:
    """smoothxg defaults: scores 1,4,6,2,26,1 (src/main.cpp:322-327), wb=311/wf=0.03 (src/smooth.cpp:266-271)."""
    return PdParams(match, mismatch, gap_open1, gap_ext1, gap_open2, gap_ext2,
                    1 if local else 0, 311 if banded else -1, 0.03, int(out_cons), int(out_msa))


@dataclass
class Dump:
    raw: np.ndarray

    def _u64(self, lo):
        return (int(self.raw[lo]) & 0xFFFFFFFF) | ((int(self.raw[lo + 1]) & 0xFFFFFFFF) << 32)

    @property
    def n_node(self): return int(self.raw[PD_N_NODE])
    @property
    def n_seq(self): return int(self.raw[PD_N_SEQ])
    @property
    def inband_cells(self): return self._u64(PD_INBAND_LO)
    @property
    def full_cells(self): return self._u64(PD_FULL_LO)
    @property
    def edge_rows(self): return self._u64(PD_EDGE_ROWS_LO)

    def sections(self) -> dict:
        r = self.raw
        n, s = self.n_node, self.n_seq
        o = PD_HEADER_LEN
        out = {}

        def take(name, k):
            nonlocal o
            out[name] = r[o:o + k]
            o += k
        take("base", n)
        take("in_n", n); take("in_id", int(r[PD_N_IN_TOT])); take("in_w", int(r[PD_N_IN_TOT]))
        take("out_n", n); take("out_id", int(r[PD_N_OUT_TOT])); take("out_w", int(r[PD_N_OUT_TOT]))
        take("aln_n", n); take("aln_id", int(r[PD_N_ALN_TOT]))
        take("path_len", s); take("path_node", int(r[PD_PATH_TOT]))
        take("cons_node", max(int(r[PD_CONS_LEN]), 0))
        take("msa", int(r[PD_MSA_ROWS]) * max(int(r[PD_MSA_LEN]), 0))
        take("best_score", s); take("n_cigar", s)
        take("cigar", 2 * int(r[PD_CIGAR_TOT]))
        assert o == r.shape[0], (o, r.shape)
        return out

    def result_part(self) -> np.ndarray:
        """Everything smoothxg consumes (graph, paths, consensus, msa); excludes the instrumentation
        tail (scores/cigars) and the cell counters, which only the instrumented drive fills."""
        r = self.raw
        n_tail = 2 * self.n_seq + 2 * int(r[PD_CIGAR_TOT])
        body = r[PD_HEADER_LEN:r.shape[0] - n_tail]
        hdr = r[:PD_CIGAR_TOT]
        return np.concatenate([hdr, body])

    def compare_part(self) -> np.ndarray:
        """Everything an instrumented run pins: result_part plus per-sequence scores, cigars and the
        in-band cell count (header slots up to PD_INBAND_HI); excludes the full-matrix / edge-row counters."""
        r = self.raw
        return np.concatenate([r[:PD_FULL_LO], r[PD_HEADER_LEN:]])


def _as(arr, dtype):
    return np.ascontiguousarray(arr, dtype=dtype)


class _Checker:
    def __init__(self, lib, fn_name, free_name):
        self.lib = lib
        self.fn = getattr(lib, fn_name)
        self.fn.restype = C.POINTER(C.c_int32)
        self.fn.argtypes = [C.POINTER(PdParams), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                            C.POINTER(C.c_int64)]
        self.free = getattr(lib, free_name)
        self.free.argtypes = [C.c_void_p]
        self.free.restype = None

    def poa_block(self, params: PdParams, seq_len, bases, weight, instrument=True) -> Dump | None:
        seq_len = _as(seq_len, np.int32); bases = _as(bases, np.uint8); weight = _as(weight, np.int32)
        n = C.c_int64(0)
        p = self.fn(C.byref(params), int(seq_len.shape[0]), seq_len.ctypes.data, bases.ctypes.data,
                    weight.ctypes.data, int(instrument), C.byref(n))
        if not p:
            return None
        arr = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
        self.free(p)
        return Dump(arr)

    def poa_batch(self, params, batch, instrument=True):
        return [self.poa_block(params, *batch.block(b), instrument=instrument) for b in range(batch.n_blocks)]


class Oracle(_Checker):
    """Scalar C restatement (oracle/poa_oracle.c)."""

    def __init__(self):
        path = os.path.join(HERE, "libpoa_oracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        lib = C.CDLL(path)
        super().__init__(lib, "oracle_poa_block", "oracle_free")
        lib.oracle_set_lane_counts.argtypes = [C.c_int, C.c_int]

    def set_lane_counts(self, pn16=32, pn32=16):
        self.lib.oracle_set_lane_counts(pn16, pn32)


def ref_available(isa: str | None = None) -> bool:
    return _ref_path(isa) is not None


def _cpu_has(flag: str) -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return flag in line.split()
    except OSError:
        pass
    return False


def _ref_path(isa):
    order = [isa] if isa else (["avx512"] if _cpu_has("avx512bw") else []) + (["avx2"] if _cpu_has("avx2") else []) + ["sse41"]
    for name in order:
        p = os.path.join(HERE, "_ref", f"libabpoa_ref_{name}.so")
        if os.path.exists(p):
            return p
    return None


class RefAbpoa(_Checker):
    """Unmodified vendored abPOA v1.5.4 driven like smooth_abpoa (oracle/ref_shim.c)."""

    def __init__(self, isa: str | None = None):
        path = _ref_path(isa)
        if path is None:
            raise FileNotFoundError("oracle/_ref/libabpoa_ref_*.so missing: run `make -C oracle ref` where /root/reference exists")
        lib = C.CDLL(path)
        super().__init__(lib, "ref_poa_block", "ref_free")
        lib.ref_simd_name.restype = C.c_char_p
        lib.ref_poa_batch_timed.restype = C.c_double
        lib.ref_poa_batch_timed.argtypes = [C.POINTER(PdParams), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        self.path = path

    @property
    def simd(self) -> str:
        return self.lib.ref_simd_name().decode()

    def batch_timed(self, params, batch, n_threads=0, want_hash=False):
        """Wall seconds for abpoa_poa over the whole batch with an OpenMP dynamic loop (src/smooth.cpp:1931)."""
        h = np.zeros(batch.n_blocks, dtype=np.uint64) if want_hash else None
        secs = self.lib.ref_poa_batch_timed(
            C.byref(params), batch.n_blocks, batch.block_seq_off.ctypes.data, batch.seq_len.ctypes.data,
            batch.seq_off.ctypes.data, batch.bases.ctypes.data, batch.weight.ctypes.data, int(n_threads),
            h.ctypes.data if want_hash else None)
        return (secs, h) if want_hash else secs
