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
from tests.golden_io import load_cases, load_real_cases, pd_params
from tests.helpers import view_to_dump

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu_poa.cpp")
OUT = os.path.join(HERE, "emu", "_build", "libpoa_emu.so")
CASES = load_cases()


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC] + [os.path.join(HERE, "..", "smoothxg_b200", "csrc", f) for f in ("poa_core.cuh", "poa_fill16.cuh", "poa_host.hpp", "poa_wire.hpp")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        inc = "/usr/local/cuda/include"
        subprocess.check_call(["/usr/bin/g++", "-O1", "-fPIC", "-shared", "-std=c++17", f"-I{inc}", "-o", OUT, SRC])
    lib = C.CDLL(OUT)
    chk = _Checker(lib, "emu_poa_block", "emu_free")
    wire = _Checker(lib, "emu_poa_block_wire", "emu_free")
    return chk, wire


OUT32 = os.path.join(HERE, "emu", "_build", "libpoa_emu32.so")
# cases the 32-lane emulation runs in the default CPU suite (the rest: POA_EMU32_ALL=1)
FAST32 = {"abpoa_seq_fa_global", "abpoa_seq_fa_local", "abpoa_test_fa", "abpoa_example_c", "edge_shapes", "edge_shapes_local",
          "syn_local", "syn_unbanded", "syn_presets", "syn_divergent", "affine_seq_fa", "linear_seq_fa", "affine_edge_shapes",
          "linear_edge_shapes", "affine_local", "linear_local"}


@pytest.fixture(scope="module")
def emu32():
    """The same device code with 32 lock-step lanes emulated as fibers (-DPOA_EMU_LANES=32): exercises the
    warp-level logic -- shuffle scans, reductions, the lane-striped packed 16-bit fill of poa_fill16.cuh."""
    os.makedirs(os.path.dirname(OUT32), exist_ok=True)
    csrc = os.path.join(HERE, "..", "smoothxg_b200", "csrc")
    deps = [SRC] + [os.path.join(csrc, f) for f in ("poa_core.cuh", "poa_fill16.cuh", "poa_host.hpp", "poa_wire.hpp")]
    if not os.path.exists(OUT32) or any(os.path.getmtime(d) > os.path.getmtime(OUT32) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O1", "-fPIC", "-shared", "-std=c++17", "-DPOA_EMU_LANES=32",
                               "-I/usr/local/cuda/include", "-o", OUT32, SRC])
    return _Checker(C.CDLL(OUT32), "emu_poa_block", "emu_free")


OUT32X4 = os.path.join(HERE, "emu", "_build", "libpoa_emu32x4.so")
FAST32X4 = {"abpoa_seq_fa_global", "abpoa_seq_fa_local", "abpoa_example_c", "edge_shapes", "edge_shapes_local", "syn_presets", "affine_seq_fa"}


@pytest.fixture(scope="module")
def emu32x4():
    """Four warps of 32 emulated lanes per POA block (-DPOA_EMU_NW=4): the multi-warp packed fill (chunks dealt to
    warps, carry chain through shared memory) and the generic multi-warp fill, cross-warp barriers included."""
    os.makedirs(os.path.dirname(OUT32X4), exist_ok=True)
    csrc = os.path.join(HERE, "..", "smoothxg_b200", "csrc")
    deps = [SRC] + [os.path.join(csrc, f) for f in ("poa_core.cuh", "poa_fill16.cuh", "poa_host.hpp", "poa_wire.hpp")]
    if not os.path.exists(OUT32X4) or any(os.path.getmtime(d) > os.path.getmtime(OUT32X4) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O1", "-fPIC", "-shared", "-std=c++17", "-DPOA_EMU_LANES=32", "-DPOA_EMU_NW=4",
                               "-I/usr/local/cuda/include", "-o", OUT32X4, SRC])
    return _Checker(C.CDLL(OUT32X4), "emu_poa_block", "emu_free")


@pytest.mark.parametrize("name,batch,p,dumps", CASES, ids=[c[0] for c in CASES])
def test_emulated_multi_warp_logic_matches_golden(emu32x4, name, batch, p, dumps):
    if name not in FAST32X4 and not os.environ.get("POA_EMU32_ALL"):
        pytest.skip("slow under the fiber emulation; set POA_EMU32_ALL=1")
    for b in range(batch.n_blocks):
        got = emu32x4.poa_block(pd_params(p), *batch.block(b))
        assert got is not None
        assert np.array_equal(got.compare_part(), dumps[b].compare_part()), f"{name} block {b}"
        assert got.edge_rows == dumps[b].edge_rows


@pytest.mark.parametrize("name,batch,p,dumps", CASES, ids=[c[0] for c in CASES])
def test_emulated_warp_logic_matches_golden(emu32, name, batch, p, dumps):
    if name not in FAST32 and not os.environ.get("POA_EMU32_ALL"):
        pytest.skip("slow under the fiber emulation; set POA_EMU32_ALL=1")
    for b in range(batch.n_blocks):
        got = emu32.poa_block(pd_params(p), *batch.block(b))
        assert got is not None
        assert np.array_equal(got.compare_part(), dumps[b].compare_part()), f"{name} block {b}"
        assert got.edge_rows == dumps[b].edge_rows


@pytest.mark.parametrize("name,batch,p,dumps", CASES, ids=[c[0] for c in CASES])
def test_emulated_device_logic_matches_golden(emu, name, batch, p, dumps):
    chk, _ = emu
    for b in range(batch.n_blocks):
        got = chk.poa_block(pd_params(p), *batch.block(b))
        assert got is not None
        assert np.array_equal(got.compare_part(), dumps[b].compare_part()), f"{name} block {b}"
        assert got.edge_rows == dumps[b].edge_rows


@pytest.mark.parametrize("mode,blocks", [("drb1_global", (0, 7, 16)), ("drb1_local", (3,))])
def test_emulated_device_logic_on_real_blocks(emu, mode, blocks):
    chk, _ = emu
    _, batch, p, dumps = next(c for c in load_real_cases() if c[0] == mode)
    for b in blocks:
        got = chk.poa_block(pd_params(p), *batch.block(b))
        assert got is not None
        assert np.array_equal(got.compare_part(), dumps[b].compare_part()), f"{mode} block {b}"


def test_emulated_warp_logic_on_a_real_block(emu32):
    _, batch, p, dumps = next(c for c in load_real_cases() if c[0] == "drb1_global")
    b = 4
    got = emu32.poa_block(pd_params(p), *batch.block(b))
    assert np.array_equal(got.compare_part(), dumps[b].compare_part())


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


def test_wide_wire_format_roundtrip(emu, oracle):
    """Total dedup weight >= 65 536 switches the result body to its 32-bit form (WIRE_WIDE, poa_core.cuh): device writer
    (emulated) -> poa_wire.hpp decoder -> views == oracle."""
    from oracle.oracle import make_params
    from smoothxg_b200 import synth
    from smoothxg_b200.shard import merge_parts
    _, wire = emu
    batch = synth.make_batch(n_blocks=2, n_seqs=24, length=120, seed=77, divergence=0.7)
    batch.weight[:] = 3000 + (np.arange(batch.weight.shape[0]) % 7) * 500
    p = make_params(out_msa=True)
    for b in range(batch.n_blocks):
        w = wire.poa_block(p, *batch.block(b)).raw
        assert w[18] == 3  # H_FORMAT == WIRE_WIDE
        hdr, arena = merge_parts(1, [(np.array([0]), w[:engine.HDR_WORDS], w[engine.HDR_WORDS:])])
        res = engine.result_from_parts(hdr, arena)
        assert np.array_equal(view_to_dump(res.block(0)).compare_part(), oracle.poa_block(p, *batch.block(b)).compare_part())
        res.close()
