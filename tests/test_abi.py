"""CPU: the C-ABI shared library loads and exports every symbol include/poa_b200.h declares; struct layouts
match the ctypes mirrors; without a GPU the engine refuses loudly (no CPU fallback exists)."""
import ctypes as C
import os
import re

import pytest

from smoothxg_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "poa_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(poa_b200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/poa_b200.h but not exported"
    assert sorted(engine.ABI_SYMBOLS) == syms
    assert lib.poa_b200_abi_version() == 1


def test_struct_layouts():
    assert C.sizeof(engine.PoaParams) == 44
    assert C.sizeof(engine.EngineOpts) == 32
    assert C.sizeof(engine.Stats) == 8 * 3 + 8 * 4 + 4 * 4 + 8 + 8 * 8
    assert engine.load_library().poa_b200_strerror(engine.EUNSUP) == b"unsupported parameters"


def test_no_cpu_fallback_without_gpu():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(engine.PoaError) as e:
        engine.PoaEngine(device=0)
    assert e.value.code == engine.ECUDA


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "smoothxg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                code = [ln for ln in txt.splitlines() if re.match(r"\s*(import|from|#include)\b", ln)]
                assert not any("oracle" in ln for ln in code), f
                assert "libpoa_oracle" not in txt and "libabpoa_ref" not in txt and "CDLL(\"oracle" not in txt, f


def test_encode_bases_matches_abpoa_table():
    """poa_b200_encode_bases == ab_nt4_table (deps/abPOA/src/abpoa_seq.c:15-32), all 256 byte values."""
    import numpy as np
    from smoothxg_b200 import synth
    allb = bytes(range(1, 256)) + b"ACGTNacgtnUu-RYKM"
    got = engine.encode_bases(allb)
    want = synth._ENC[np.frombuffer(allb, dtype=np.uint8)]
    assert np.array_equal(got, want)
