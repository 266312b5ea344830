"""ctypes binding of the C ABI in include/poa_b200.h (libpoa_b200.so, hand-written sm_100a CUDA).

Python is only the harness here (tests, bench, multi-GPU driver); the product is the shared library.
The names mirror the reference's per-block call site: `PoaParams` carries what smooth_abpoa puts into
abpoa_para_t (reference src/smooth.cpp:256-297), `PoaEngine.run_batch` stands where the OpenMP loop
calls abpoa_poa (src/smooth.cpp:337, :1931), and `BlockView` exposes the abpoa_t fields that
build_odgi_abPOA and the MAF code read afterwards (src/smooth.cpp:362-516, :2442-2574).

There is deliberately no fallback: if the library is missing or no Blackwell GPU is present this
module raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libpoa_b200.so")

# keep in sync with include/poa_b200.h
OK, ESLAB, EARENA, EINTERNAL, EUNSUP, EBLOCK, ECUDA, EARG, ENOMEM = range(9)
HDR_WORDS = 20  # POA_B200_HDR_WORDS
H_OFF_LO, H_OFF_HI = 11, 12  # header slots holding a block body's word offset in its arena

ABI_SYMBOLS = [
    "poa_b200_abi_version", "poa_b200_strerror", "poa_b200_last_error",
    "poa_b200_encode_bases", "poa_b200_engine_create", "poa_b200_engine_destroy", "poa_b200_engine_trim",
    "poa_b200_run_batch", "poa_b200_poa_block", "poa_b200_submit_block", "poa_b200_wait_block",
    "poa_b200_batch_upload", "poa_b200_batch_launch", "poa_b200_batch_download", "poa_b200_batch_finish",
    "poa_b200_batch_free", "poa_b200_batch_stats", "poa_b200_batch_device_result", "poa_b200_result_from_parts", "poa_b200_result_from_device_parts",
    "poa_b200_result_n_blocks", "poa_b200_result_block", "poa_b200_result_release_block", "poa_b200_result_block_hash", "poa_b200_result_stats", "poa_b200_result_free",
    "poa_b200_block_graph", "poa_b200_graph_view", "poa_b200_graph_free", "poa_b200_block_final_graph", "poa_b200_final_graph_view",
]


class PoaParams(C.Structure):
    """poa_b200_params_t; smoothxg defaults: scores 1,4,6,2,26,1 (src/main.cpp:322-327), wb=311, wf=0.03."""
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open1", C.c_int32),
                ("gap_ext1", C.c_int32), ("gap_open2", C.c_int32), ("gap_ext2", C.c_int32),
                ("align_mode", C.c_int32), ("wb", C.c_int32), ("wf", C.c_float),
                ("out_cons", C.c_int32), ("out_msa", C.c_int32)]


def make_params(match=1, mismatch=4, gap_open1=6, gap_ext1=2, gap_open2=26, gap_ext2=1,
                local=False, banded=True, out_cons=True, out_msa=False) -> PoaParams:
    return PoaParams(match, mismatch, gap_open1, gap_ext1, gap_open2, gap_ext2,
                     1 if local else 0, 311 if banded else -1, 0.03, int(out_cons), int(out_msa))


class EngineOpts(C.Structure):
    _fields_ = [("warps_per_block", C.c_int32), ("ctas_per_sm", C.c_int32), ("emit_cigar", C.c_int32),
                ("flags", C.c_int32), ("slab_rows_factor", C.c_double), ("device_mem_budget", C.c_int64)]


class _BlockView(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_node", C.c_int32), ("n_seq", C.c_int32), ("cons_len", C.c_int32),
                ("msa_len", C.c_int32), ("msa_rows", C.c_int32),
                ("base", C.POINTER(C.c_int32)), ("in_n", C.POINTER(C.c_int32)), ("in_id", C.POINTER(C.c_int32)),
                ("in_w", C.POINTER(C.c_int32)), ("out_n", C.POINTER(C.c_int32)), ("out_id", C.POINTER(C.c_int32)),
                ("out_w", C.POINTER(C.c_int32)), ("aln_n", C.POINTER(C.c_int32)), ("aln_id", C.POINTER(C.c_int32)),
                ("path_len", C.POINTER(C.c_int32)), ("path_node", C.POINTER(C.c_int32)),
                ("cons_node", C.POINTER(C.c_int32)), ("msa", C.POINTER(C.c_uint8)),
                ("best_score", C.POINTER(C.c_int32)), ("n_cigar", C.POINTER(C.c_int32)),
                ("cigar", C.POINTER(C.c_uint64)),
                ("in_total", C.c_int64), ("out_total", C.c_int64), ("aln_total", C.c_int64),
                ("path_total", C.c_int64), ("cigar_total", C.c_int64), ("inband_cells", C.c_int64)]


class _GraphView(C.Structure):
    _fields_ = [("n_node", C.c_int32), ("node_id", C.POINTER(C.c_int32)), ("node_base", C.POINTER(C.c_char)),
                ("n_edge", C.c_int32), ("edge_from", C.POINTER(C.c_int32)), ("edge_to", C.POINTER(C.c_int32)),
                ("n_path", C.c_int32), ("path_off", C.POINTER(C.c_int64)), ("path_node", C.POINTER(C.c_int32))]


@dataclass
class BlockGraph:
    """poa_b200_graph_view_t: what build_odgi_abPOA leaves in the odgi graph (reference src/smooth.cpp:2442-2574)."""
    node_id: np.ndarray
    node_base: bytes
    edge_from: np.ndarray
    edge_to: np.ndarray
    path_off: np.ndarray
    path_node: np.ndarray

    def path(self, i: int) -> np.ndarray:
        return self.path_node[self.path_off[i]:self.path_off[i + 1]]


class _FinalGraphView(C.Structure):
    _fields_ = [("n_node", C.c_int32), ("seq_off", C.POINTER(C.c_int64)), ("seq", C.POINTER(C.c_char)),
                ("n_edge", C.c_int32), ("edge_from", C.POINTER(C.c_int32)), ("edge_to", C.POINTER(C.c_int32)),
                ("n_path", C.c_int32), ("path_off", C.POINTER(C.c_int64)), ("path_node", C.POINTER(C.c_int32))]


@dataclass
class FinalGraph:
    """poa_b200_final_graph_view_t: the block graph smooth_abpoa returns (reference src/smooth.cpp:545-620), ids 1..n."""
    node_seq: list
    edge_from: np.ndarray
    edge_to: np.ndarray
    path_off: np.ndarray
    path_node: np.ndarray

    def path(self, i: int) -> np.ndarray:
        return self.path_node[self.path_off[i]:self.path_off[i + 1]]


class Stats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
                ("inband_cells", C.c_int64), ("edge_row_cells", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("kernel_launches", C.c_int32), ("retried_blocks", C.c_int32), ("n_ctas", C.c_int32),
                ("warps_per_block", C.c_int32), ("workspace_bytes", C.c_int64), ("phase_cycles", C.c_int64 * 8)]

    def as_dict(self):
        names = ["rows", "fill", "backtrack", "fuse", "toposort", "finalize", "total", "spare"]
        d = {f: getattr(self, f) for f, _ in self._fields_ if f != "phase_cycles"}
        d["phase_cycles"] = {n: int(self.phase_cycles[i]) for i, n in enumerate(names)}
        return d


class PoaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"poa_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load_library() -> C.CDLL:
    """Load libpoa_b200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("POA_B200_LIB", LIB_PATH)  # alternative builds of the same CUDA library (tuning experiments)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(nvcc -gencode arch=compute_100a,code=sm_100a); there is no CPU fallback")
    lib = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.poa_b200_abi_version.restype = C.c_int
    lib.poa_b200_strerror.restype = C.c_char_p
    lib.poa_b200_strerror.argtypes = [C.c_int]
    lib.poa_b200_last_error.restype = C.c_char_p
    lib.poa_b200_encode_bases.argtypes = [C.c_char_p, i64, vp]
    lib.poa_b200_encode_bases.restype = None
    lib.poa_b200_engine_create.argtypes = [C.c_int, C.POINTER(EngineOpts), C.POINTER(vp)]
    lib.poa_b200_engine_destroy.argtypes = [vp]
    lib.poa_b200_engine_destroy.restype = None
    lib.poa_b200_engine_trim.argtypes = [vp]
    batch_args = [vp, C.POINTER(PoaParams), i64, vp, vp, vp, vp, vp, C.POINTER(vp)]
    lib.poa_b200_run_batch.argtypes = batch_args
    lib.poa_b200_batch_upload.argtypes = batch_args
    lib.poa_b200_poa_block.argtypes = [vp, C.POINTER(PoaParams), i32, C.POINTER(C.c_void_p), vp, vp, C.POINTER(vp)]
    lib.poa_b200_submit_block.argtypes = [vp, C.POINTER(PoaParams), i32, C.POINTER(C.c_void_p), vp, vp, C.POINTER(C.c_uint64)]
    lib.poa_b200_wait_block.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
    lib.poa_b200_batch_launch.argtypes = [vp, vp]
    lib.poa_b200_batch_download.argtypes = [vp, vp, C.POINTER(vp)]
    lib.poa_b200_batch_finish.argtypes = [vp, vp]
    lib.poa_b200_batch_free.argtypes = [vp]
    lib.poa_b200_batch_free.restype = None
    lib.poa_b200_batch_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.poa_b200_batch_device_result.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), C.POINTER(i32), C.POINTER(vp)]
    lib.poa_b200_result_from_parts.argtypes = [i64, vp, vp, i64, C.POINTER(vp)]
    lib.poa_b200_result_from_device_parts.argtypes = [vp, i64, vp, vp, i64, vp, C.POINTER(vp)]
    lib.poa_b200_result_n_blocks.argtypes = [vp]
    lib.poa_b200_result_n_blocks.restype = i64
    lib.poa_b200_result_block.argtypes = [vp, i64, C.POINTER(_BlockView)]
    lib.poa_b200_result_block_hash.argtypes = [vp, i64, C.POINTER(C.c_uint64)]
    lib.poa_b200_result_release_block.argtypes = [vp, i64]
    lib.poa_b200_result_release_block.restype = None
    lib.poa_b200_block_graph.argtypes = [C.POINTER(_BlockView), i32, i32, C.POINTER(vp)]
    lib.poa_b200_graph_view.argtypes = [vp, C.POINTER(_GraphView)]
    lib.poa_b200_block_final_graph.argtypes = [C.POINTER(_BlockView), i32, i32, C.POINTER(vp)]
    lib.poa_b200_final_graph_view.argtypes = [vp, C.POINTER(_FinalGraphView)]
    lib.poa_b200_graph_free.argtypes = [vp]
    lib.poa_b200_graph_free.restype = None
    lib.poa_b200_result_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.poa_b200_result_free.argtypes = [vp]
    lib.poa_b200_result_free.restype = None
    _lib = lib
    return lib


def encode_bases(ascii_seq: bytes | str) -> np.ndarray:
    """ASCII -> abPOA codes through the library (poa_b200_encode_bases; reference src/smooth.cpp:304-313)."""
    if isinstance(ascii_seq, str):
        ascii_seq = ascii_seq.encode()
    out = np.empty(len(ascii_seq), dtype=np.uint8)
    load_library().poa_b200_encode_bases(ascii_seq, len(ascii_seq), out.ctypes.data)
    return out


def _check(lib, rc, allow=()):
    if rc != OK and rc not in allow:
        raise PoaError(rc, f"{lib.poa_b200_strerror(rc).decode()}: {lib.poa_b200_last_error().decode()}")
    return rc


def _arr(ptr, n, dtype=np.int32):
    if n <= 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).copy()


@dataclass
class BlockView:
    """One finished POA block (copies of the flat arrays of poa_b200_block_view_t)."""
    status: int
    n_node: int
    n_seq: int
    cons_len: int
    msa_len: int
    msa_rows: int
    base: np.ndarray
    in_n: np.ndarray
    in_id: np.ndarray
    in_w: np.ndarray
    out_n: np.ndarray
    out_id: np.ndarray
    out_w: np.ndarray
    aln_n: np.ndarray
    aln_id: np.ndarray
    path_len: np.ndarray
    path_node: np.ndarray
    cons_node: np.ndarray
    msa: np.ndarray
    best_score: np.ndarray
    n_cigar: np.ndarray
    cigar: np.ndarray
    inband_cells: int

    def consensus(self) -> str:
        return "".join("ACGTN"[int(self.base[i])] for i in self.cons_node)


class PoaResult:
    def __init__(self, lib, handle):
        self._lib, self._h = lib, handle

    def __len__(self):
        return int(self._lib.poa_b200_result_n_blocks(self._h))

    def block(self, i: int) -> BlockView:
        v = _BlockView()
        _check(self._lib, self._lib.poa_b200_result_block(self._h, i, C.byref(v)))
        if v.status != OK:
            z = np.zeros(0, dtype=np.int32)
            return BlockView(v.status, 0, v.n_seq, -1, -1, 0, z, z, z, z, z, z, z, z, z, z, z, z,
                             np.zeros(0, np.uint8), z, z, np.zeros(0, np.uint64), 0)
        n, s = v.n_node, v.n_seq
        try:
            return self._copy_view(v, n, s)
        finally:
            self._lib.poa_b200_result_release_block(self._h, i)  # everything was copied: drop the library's flat expansion

    @staticmethod
    def _copy_view(v, n, s) -> BlockView:
        msa = _arr(v.msa, v.msa_rows * max(v.msa_len, 0), np.uint8)
        cig = np.zeros(0, dtype=np.uint64)
        if v.cigar_total > 0:
            w = np.ctypeslib.as_array(C.cast(v.cigar, C.POINTER(C.c_uint32)), shape=(2 * int(v.cigar_total),)).copy()
            cig = w[0::2].astype(np.uint64) | (w[1::2].astype(np.uint64) << np.uint64(32))
        return BlockView(v.status, n, s, v.cons_len, v.msa_len, v.msa_rows,
                         _arr(v.base, n), _arr(v.in_n, n), _arr(v.in_id, v.in_total), _arr(v.in_w, v.in_total),
                         _arr(v.out_n, n), _arr(v.out_id, v.out_total), _arr(v.out_w, v.out_total),
                         _arr(v.aln_n, n), _arr(v.aln_id, v.aln_total),
                         _arr(v.path_len, s), _arr(v.path_node, v.path_total), _arr(v.cons_node, max(v.cons_len, 0)),
                         msa, _arr(v.best_score, s), _arr(v.n_cigar, s), cig, int(v.inband_cells))

    def block_hash(self, i: int) -> int:
        """FNV-1a of block i's graph (poa_b200_result_block_hash): comparable with oracle/ref_shim.c's per-block hash."""
        h = C.c_uint64()
        _check(self._lib, self._lib.poa_b200_result_block_hash(self._h, i, C.byref(h)))
        self._lib.poa_b200_result_release_block(self._h, i)
        return int(h.value)

    def block_graph(self, i: int, padding_len: int = 0, include_consensus: bool = True) -> BlockGraph:
        """The per-block graph smoothxg's build_odgi_abPOA would leave behind (poa_b200_block_graph)."""
        v = _BlockView()
        _check(self._lib, self._lib.poa_b200_result_block(self._h, i, C.byref(v)))
        g = C.c_void_p()
        _check(self._lib, self._lib.poa_b200_block_graph(C.byref(v), padding_len, int(include_consensus), C.byref(g)))
        gv = _GraphView()
        _check(self._lib, self._lib.poa_b200_graph_view(g, C.byref(gv)))
        off = _arr(gv.path_off, gv.n_path + 1, np.int64)
        out = BlockGraph(_arr(gv.node_id, gv.n_node), bytes(gv.node_base[:gv.n_node]) if gv.n_node else b"",
                         _arr(gv.edge_from, gv.n_edge), _arr(gv.edge_to, gv.n_edge), off,
                         _arr(gv.path_node, int(off[-1]) if off.size else 0))
        self._lib.poa_b200_graph_free(g)
        self._lib.poa_b200_result_release_block(self._h, i)
        return out

    def final_graph(self, i: int, padding_len: int = 0, include_consensus: bool = True) -> FinalGraph:
        """The graph smooth_abpoa returns for block i: unchopped, topologically ordered, compact ids (poa_b200_block_final_graph)."""
        v = _BlockView()
        _check(self._lib, self._lib.poa_b200_result_block(self._h, i, C.byref(v)))
        g = C.c_void_p()
        _check(self._lib, self._lib.poa_b200_block_final_graph(C.byref(v), padding_len, int(include_consensus), C.byref(g)))
        gv = _FinalGraphView()
        _check(self._lib, self._lib.poa_b200_final_graph_view(g, C.byref(gv)))
        so = _arr(gv.seq_off, gv.n_node + 1, np.int64)
        seq = bytes(gv.seq[:int(so[-1])]) if gv.n_node else b""
        off = _arr(gv.path_off, gv.n_path + 1, np.int64)
        out = FinalGraph([seq[int(so[k]):int(so[k + 1])].decode() for k in range(gv.n_node)],
                         _arr(gv.edge_from, gv.n_edge), _arr(gv.edge_to, gv.n_edge), off, _arr(gv.path_node, int(off[-1]) if off.size else 0))
        self._lib.poa_b200_graph_free(g)
        self._lib.poa_b200_result_release_block(self._h, i)
        return out

    def stats(self) -> dict:
        s = Stats()
        _check(self._lib, self._lib.poa_b200_result_stats(self._h, C.byref(s)))
        return s.as_dict()

    def close(self):
        if self._h:
            self._lib.poa_b200_result_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


def result_from_parts(hdr: np.ndarray, arena: np.ndarray) -> PoaResult:
    """Host result from gathered header / arena words (poa_b200_result_from_parts)."""
    lib = load_library()
    hdr = np.ascontiguousarray(hdr, dtype=np.int32); arena = np.ascontiguousarray(arena, dtype=np.int32)
    r = C.c_void_p()
    _check(lib, lib.poa_b200_result_from_parts(hdr.shape[0] // HDR_WORDS, hdr.ctypes.data, arena.ctypes.data, arena.shape[0], C.byref(r)))
    return PoaResult(lib, r)


class DeviceBatch:
    """A batch whose inputs are resident in HBM (staged API)."""

    def __init__(self, lib, handle, keep):
        self._lib, self._h, self._keep = lib, handle, keep

    def launch(self, stream: int | None = None):
        _check(self._lib, self._lib.poa_b200_batch_launch(self._h, C.c_void_p(stream or 0)))

    def finish(self, stream: int | None = None):
        _check(self._lib, self._lib.poa_b200_batch_finish(self._h, C.c_void_p(stream or 0)))

    def download(self, stream: int | None = None, allow_block_errors=False) -> PoaResult:
        r = C.c_void_p()
        rc = self._lib.poa_b200_batch_download(self._h, C.c_void_p(stream or 0), C.byref(r))
        _check(self._lib, rc, allow=(EBLOCK,) if allow_block_errors else ())
        return PoaResult(self._lib, r)

    def stats(self) -> dict:
        s = Stats()
        _check(self._lib, self._lib.poa_b200_batch_stats(self._h, C.byref(s)))
        return s.as_dict()

    def device_result(self, n_blocks: int):
        """(d_hdr_ptr, [(d_arena_ptr, words), ...], block_arena[n_blocks]) of a finished batch: raw device pointers
        for the multi-GPU gather (smoothxg_b200/shard.py wraps them as torch tensors)."""
        hdr, ar, bl = C.c_void_p(), C.c_void_p(), C.c_void_p()
        words, n_ar = C.c_int64(), C.c_int32()
        _check(self._lib, self._lib.poa_b200_batch_device_result(self._h, 0, C.byref(hdr), C.byref(ar), C.byref(words), C.byref(n_ar), C.byref(bl)))
        arenas = []
        for i in range(n_ar.value):
            _check(self._lib, self._lib.poa_b200_batch_device_result(self._h, i, None, C.byref(ar), C.byref(words), None, None))
            arenas.append((ar.value, int(words.value)))
        block_arena = np.ctypeslib.as_array(C.cast(bl, C.POINTER(C.c_int32)), shape=(n_blocks,)).copy() if n_blocks else np.zeros(0, np.int32)
        return hdr.value, arenas, block_arena

    def close(self):
        if self._h:
            self._lib.poa_b200_batch_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class PoaEngine:
    """One engine per GPU (one process per GPU in the multi-GPU driver)."""

    def __init__(self, device: int = 0, warps_per_block: int = 0, ctas_per_sm: int = 0, emit_cigar: bool = False,
                 slab_rows_factor: float = 0.0, device_mem_budget: int = 0, flags: int = 0):
        self._h = None
        self._lib = load_library()
        opts = EngineOpts(warps_per_block, ctas_per_sm, int(emit_cigar), flags, slab_rows_factor, device_mem_budget)
        h = C.c_void_p()
        _check(self._lib, self._lib.poa_b200_engine_create(device, C.byref(opts), C.byref(h)))
        self._h = h

    def _args(self, batch):
        bso = _c(batch.block_seq_off, np.int64); sl = _c(batch.seq_len, np.int32); so = _c(batch.seq_off, np.int64)
        ba = _c(batch.bases, np.uint8); wt = _c(batch.weight, np.int32)
        keep = (bso, sl, so, ba, wt)
        return keep, (int(bso.shape[0] - 1), bso.ctypes.data, sl.ctypes.data, so.ctypes.data, ba.ctypes.data, wt.ctypes.data)

    def run_batch(self, batch, params: PoaParams, allow_block_errors=False) -> PoaResult:
        """Host buffers in, host result out: H2D, kernels, D2H (poa_b200_run_batch)."""
        keep, a = self._args(batch)
        r = C.c_void_p()
        rc = self._lib.poa_b200_run_batch(self._h, C.byref(params), *a, C.byref(r))
        _check(self._lib, rc, allow=(EBLOCK,) if allow_block_errors else ())
        del keep
        return PoaResult(self._lib, r)

    def upload(self, batch, params: PoaParams) -> DeviceBatch:
        keep, a = self._args(batch)
        h = C.c_void_p()
        _check(self._lib, self._lib.poa_b200_batch_upload(self._h, C.byref(params), *a, C.byref(h)))
        return DeviceBatch(self._lib, h, keep)

    def result_from_device(self, hdr: np.ndarray, d_arena_ptr: int, arena_words: int, stream: int | None = None) -> PoaResult:
        """Host result from host headers + arena words resident on this engine's GPU (poa_b200_result_from_device_parts)."""
        hdr = np.ascontiguousarray(hdr, dtype=np.int32)
        r = C.c_void_p()
        _check(self._lib, self._lib.poa_b200_result_from_device_parts(self._h, hdr.shape[0] // HDR_WORDS, hdr.ctypes.data, C.c_void_p(d_arena_ptr),
                                                                       int(arena_words), C.c_void_p(stream or 0), C.byref(r)))
        return PoaResult(self._lib, r)

    def poa_block(self, seqs, weights, params: PoaParams) -> PoaResult:
        """abpoa_poa-shaped convenience call for one block (poa_b200_poa_block)."""
        seqs = [_c(s, np.uint8) for s in seqs]
        n = len(seqs)
        ptrs = (C.c_void_p * max(n, 1))(*[s.ctypes.data for s in seqs])
        lens = _c([s.shape[0] for s in seqs], np.int32); wts = _c(weights, np.int32)
        r = C.c_void_p()
        _check(self._lib, self._lib.poa_b200_poa_block(self._h, C.byref(params), n, ptrs, lens.ctypes.data, wts.ctypes.data, C.byref(r)))
        return PoaResult(self._lib, r)

    def trim(self):
        """Return pooled device / pinned buffers to the driver (poa_b200_engine_trim)."""
        _check(self._lib, self._lib.poa_b200_engine_trim(self._h))

    def submit_block(self, seqs, weights, params: PoaParams) -> int:
        """Queue one block for the engine's coalescing dispatcher; returns a ticket (poa_b200_submit_block)."""
        seqs = [_c(s, np.uint8) for s in seqs]
        n = len(seqs)
        ptrs = (C.c_void_p * max(n, 1))(*[s.ctypes.data for s in seqs])
        lens = _c([s.shape[0] for s in seqs], np.int32); wts = _c(weights, np.int32)
        t = C.c_uint64()
        _check(self._lib, self._lib.poa_b200_submit_block(self._h, C.byref(params), n, ptrs, lens.ctypes.data, wts.ctypes.data, C.byref(t)))
        return int(t.value)

    def wait_block(self, ticket: int) -> PoaResult:
        r = C.c_void_p()
        _check(self._lib, self._lib.poa_b200_wait_block(self._h, C.c_uint64(ticket), C.byref(r)))
        return PoaResult(self._lib, r)

    def close(self):
        if self._h:
            self._lib.poa_b200_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()
