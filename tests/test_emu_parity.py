"""CPU: the device code (smoothxg_b200/csrc/poa_core.cuh) compiled as plain C++ with one emulated thread
(tests/emu/emu_poa.cpp, -DPOA_HOST_EMU) against the golden vectors.  This checks the serial device logic
-- vector indexing, band bookkeeping, traceback, fusion, topological sort, output packing -- and the C
ABI's result accessors without a GPU.  It is a debug harness, not a product path: the warp/block
collectives are identities here and are only exercised by the -m gpu tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle.oracle import _Checker
from smoothxg_b200 import engine
from tests.golden_io import load_cases, pd_params
from tests.helpers import view_to_dump

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu_poa.cpp")
OUT = os.path.join(HERE, "emu", "_build", "libpoa_emu.so")
CASES = load_cases()


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC, os.path.join(HERE, "..", "smoothxg_b200", "csrc", "poa_core.cuh"), os.path.join(HERE, "..", "smoothxg_b200", "csrc", "poa_host.hpp")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        inc = "/usr/local/cuda/include"
        subprocess.check_call(["/usr/bin/g++", "-O1", "-fPIC", "-shared", "-std=c++17", f"-I{inc}", "-o", OUT, SRC])
    lib = C.CDLL(OUT)
    chk = _Checker(lib, "emu_poa_block", "emu_free")
    wire = _Checker(lib, "emu_poa_block_wire", "emu_free")
    return chk, wire


@pytest.mark.parametrize("name,batch,p,dumps", CASES, ids=[c[0] for c in CASES])
def test_emulated_device_logic_matches_golden(emu, name, batch, p, dumps):
    chk, _ = emu
    for b in range(batch.n_blocks):
        got = chk.poa_block(pd_params(p), *batch.block(b))
        assert got is not None
        assert np.array_equal(got.compare_part(), dumps[b].compare_part()), f"{name} block {b}"
        assert got.edge_rows == dumps[b].edge_rows


def test_result_accessors_on_wire_format(emu):
    """wire format -> poa_b200_result_from_parts -> block views -> canonical dump == golden."""
    _, wire = emu
    from smoothxg_b200.shard import merge_parts
    for name in ("abpoa_heter_fa_global", "edge_shapes", "syn_local"):
        _, batch, p, dumps = next(c for c in CASES if c[0] == name)
        parts = []
        for b in range(batch.n_blocks):
            w = wire.poa_block(pd_params(p), *batch.block(b)).raw
            parts.append((np.array([b]), w[:engine.HDR_WORDS], w[engine.HDR_WORDS:]))
        hdr, arena = merge_parts(batch.n_blocks, parts)
        res = engine.result_from_parts(hdr, arena)
        assert len(res) == batch.n_blocks
        for b in range(batch.n_blocks):
            assert np.array_equal(view_to_dump(res.block(b)).compare_part(), dumps[b].compare_part()), f"{name} block {b}"
        res.close()


def test_result_from_parts_rejects_truncated_arena(emu):
    _, wire = emu
    _, batch, p, _ = next(c for c in CASES if c[0] == "syn_global_band")
    w = wire.poa_block(pd_params(p), *batch.block(0)).raw
    with pytest.raises(engine.PoaError):
        engine.result_from_parts(w[:engine.HDR_WORDS], w[engine.HDR_WORDS:-10])
