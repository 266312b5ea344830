"""Identity estimate behind --adaptive-poa-params (reference src/smooth.cpp:1982-2062, deps/mkmh): the oracle restatement
against golden vectors of the unmodified reference headers, the device logic replayed on the host, the ABI, and (-m gpu)
the CUDA kernels against the oracle -- bit-exact hash lists, merge-match counts, identities and thresholds."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.mash_cases import make_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "mash_golden.npz")


def golden_cases():
    z = np.load(GOLDEN)
    for name in z["names"].tobytes().decode().split("\n"):
        n = int(z[f"{name}/n_seq"])
        raw = z[f"{name}/seqs"].tobytes().decode()
        seqs = raw.split("\n") if n else []
        assert len(seqs) == n
        yield name, int(z[f"{name}/kmer"]), seqs, int(z[f"{name}/kept"]), np.float32(z[f"{name}/threshold"]), z[f"{name}/pair_identity"], z[f"{name}/hashes0"]


@pytest.fixture(scope="module")
def mash_oracle():
    from oracle.mash import MashOracle
    return MashOracle()


def test_oracle_matches_reference_golden(mash_oracle):
    n = 0
    for name, k, seqs, kept, thr, ident, h0 in golden_cases():
        ko, to, io, _ = mash_oracle.block(seqs, k)
        assert ko == kept, name
        assert (np.float32(-1.0) if to is None else to) == thr, name
        assert np.array_equal(io, ident), name
        if seqs and len(seqs[0]) > k:
            assert np.array_equal(mash_oracle.hashes(seqs[0], k), h0), name
        n += 1
    assert n >= 30


def test_golden_cases_are_reproducible():
    """The committed vectors hold their own inputs; the generator's cases must still be the same strings."""
    want = {name: seqs for name, _, seqs, *_ in golden_cases()}
    for name, _, seqs in make_cases():
        assert want[name] == seqs, name


def test_oracle_matches_unmodified_reference_on_fresh_inputs(mash_oracle):
    from oracle import mash
    if not mash.ref_available():
        pytest.skip("oracle/_ref/libmash_ref.so not built (needs /root/reference)")
    ref = mash.MashRef()
    for name, k, seqs in make_cases(seed=977, n=16):
        ko, to, io, _ = mash_oracle.block(seqs, k)
        kr, tr, ir, _ = ref.block(seqs, k)
        assert ko == kr and to == tr and np.array_equal(io, ir), name
        for s in seqs[:2]:
            if len(s) > k:
                assert np.array_equal(mash_oracle.hashes(s, k), ref.hashes(s, k)), name


def test_presets_follow_the_reference_thresholds(mash_oracle):
    from smoothxg_b200 import adaptive
    # src/smooth.cpp:2026-2062: a float compared with double literals -- 0.95f and 0.9f are below 0.95 and 0.9
    for t, want in [(1.0, (1, 19, 39, 3, 81, 1)), (0.99, (1, 19, 39, 3, 81, 1)), (0.985, (1, 13, 31, 3, 51, 1)), (0.98, (1, 13, 31, 3, 51, 1)),
                    (0.97, (1, 9, 16, 2, 41, 1)), (0.96, (1, 7, 11, 2, 33, 1)), (0.95, (1, 4, 6, 2, 26, 1)), (0.9500001, (1, 7, 11, 2, 33, 1)),
                    (0.9, None), (0.9000001, (1, 4, 6, 2, 26, 1)), (0.7, None), (-1.0, None)]:
        assert adaptive.preset(t) == want, t
        assert mash_oracle.preset(t) == want, t


def _emu_lib():
    src = os.path.join(ROOT, "tests", "emu", "emu_mash.cpp")
    out = os.path.join(ROOT, "tests", "emu", "_build", "libemu_mash.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [src, os.path.join(ROOT, "smoothxg_b200", "csrc", "mash_core.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas", src, "-o", out])
    return C.CDLL(out)


def test_device_logic_replayed_on_the_host(mash_oracle):
    """mash_core.cuh compiled for the CPU: k-mer hashes, the padding-free bitonic network, pair decoding and the
    per-element match rule give the oracle's sorted lists and merge-match counts."""
    lib = _emu_lib()
    for name, k, seqs, kept, thr, ident, h0 in golden_cases():
        if seqs and len(seqs[0]) > k:
            b = seqs[0].encode()
            out = np.zeros(len(b), dtype=np.uint64)
            n = lib.emu_mash_hashes(C.c_char_p(b), len(b), k, out.ctypes.data_as(C.c_void_p))
            assert np.array_equal(out[:n], h0), name
        keep = [s for s in seqs if len(s) >= 8 * k]
        if len(keep) < 2:
            continue
        _, _, _, common = mash_oracle.block(keep, k)
        bufs = [s.encode() for s in keep]
        arr = (C.c_char_p * len(bufs))(*bufs)
        lens = (C.c_int * len(bufs))(*[len(b) for b in bufs])
        got = np.zeros(len(common), dtype=np.uint32)
        lib.emu_mash_block_common(len(bufs), arr, lens, k, got.ctypes.data_as(C.c_void_p))
        assert np.array_equal(got.astype(np.uint64), common), name


def test_abi_exports_and_host_entry_points():
    from smoothxg_b200 import adaptive, engine
    lib = engine.load_library()
    hdr = open(os.path.join(ROOT, "include", "mash_b200.h")).read()
    import re
    for sym in sorted(set(re.findall(r"\b(mash_b200_[a-z_]+)\s*\(", hdr))):
        assert hasattr(lib, sym), sym
    fb = adaptive.flatten([["A" * 200, "C" * 136, "G" * 135], ["ACGT" * 50], [], ["A" * 300, "C" * 300, "G" * 300, "T" * 10]])
    assert adaptive.pair_offsets(fb, 17).tolist() == [0, 1, 1, 1, 4]
    assert adaptive.pair_offsets(fb, 11).tolist() == [0, 3, 3, 3, 6]


def test_argument_errors_are_reported_not_fatal():
    """The reference exits the process on errors of this path; the ABI returns MASH_B200_EARG with a message."""
    from smoothxg_b200 import adaptive, engine
    lib = engine.load_library()
    adaptive.pair_offsets(adaptive.flatten([["A" * 200]]))  # binds the argtypes
    fb = adaptive.flatten([["ACGT" * 100, "ACGT" * 90]])
    thr = np.zeros(1, dtype=np.float32)
    args = lambda k, nb=1, bases=fb.bases.ctypes.data: (0, k, nb, fb.block_seq_off.ctypes.data, fb.seq_len.ctypes.data, fb.seq_off.ctypes.data,
                                                        bases, thr.ctypes.data, None, None, None, None)
    for bad in (args(0), args(33), args(17, -1), args(17, 1, None)):
        assert lib.mash_b200_block_identity(*bad) == 7
        assert lib.mash_b200_last_error()
    st = adaptive.MashStats()
    assert lib.mash_b200_block_identity(0, 17, 0, None, None, None, None, None, None, None, None, C.byref(st)) == 0 and st.n_pairs == 0  # empty batch


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from smoothxg_b200 import adaptive
    with pytest.raises(RuntimeError):
        adaptive.block_identity(adaptive.flatten([["ACGT" * 100, "ACGT" * 90]]))


# ------------------------------------------------------------------------------------------------ GPU
def _check_against_oracle(mash_oracle, blocks, kmers):
    from smoothxg_b200 import adaptive
    for k in sorted(set(kmers)):
        sel = [b for b, kk in zip(blocks, kmers) if kk == k]
        res = adaptive.block_identity(adaptive.flatten(sel), kmer=k, want_pairs=True)
        for bi, seqs in enumerate(sel):
            ko, to, io, co = mash_oracle.block(seqs, k)
            assert res["n_kept"][bi] == ko
            assert res["threshold"][bi] == (np.float32(-1.0) if to is None else to)
            lo, hi = int(res["pair_off"][bi]), int(res["pair_off"][bi + 1])
            assert hi - lo == len(io)
            assert np.array_equal(res["pair_common"][lo:hi].astype(np.uint64), co)
            assert np.array_equal(res["pair_identity"][lo:hi], io)
    return res


@pytest.mark.gpu
def test_gpu_matches_golden_and_oracle(mash_oracle):
    from smoothxg_b200 import adaptive
    cases = list(golden_cases())
    for k in sorted({c[1] for c in cases}):
        sel = [c for c in cases if c[1] == k]
        res = adaptive.block_identity(adaptive.flatten([c[2] for c in sel]), kmer=k, want_pairs=True)
        for bi, (name, _, seqs, kept, thr, ident, _) in enumerate(sel):
            assert res["n_kept"][bi] == kept, name
            assert res["threshold"][bi] == thr, name
            lo, hi = int(res["pair_off"][bi]), int(res["pair_off"][bi + 1])
            assert np.array_equal(res["pair_identity"][lo:hi], ident), name
    _check_against_oracle(mash_oracle, [c[2] for c in cases], [c[1] for c in cases])


@pytest.mark.gpu
def test_gpu_fresh_seeds_chunked_and_long_lists(mash_oracle, monkeypatch):
    cases = make_cases(seed=4242, n=20)
    rng = np.random.default_rng(9)
    long_a = "".join(rng.choice(list("ACGT"), 20000))   # > 16 K hashes: the global-memory sorting path
    long_b = long_a[:9000] + "".join(rng.choice(list("ACGT"), 500)) + long_a[9000:]
    blocks = [c[2] for c in cases] + [[long_a, long_b, long_a[3000:19000]]]
    kmers = [c[1] for c in cases] + [17]
    _check_against_oracle(mash_oracle, blocks, kmers)
    monkeypatch.setenv("MASH_B200_CHUNK_HASHES", "5000")  # several device passes per call
    res = _check_against_oracle(mash_oracle, blocks, kmers)
    assert res["stats"]["n_chunks"] >= 1


@pytest.mark.gpu
def test_gpu_benchmark_shape_properties(mash_oracle):
    """BASELINE configs[2] shape (32 x 2 kb blocks at 2 % divergence): thresholds are in [0.7, 1], identical strings
    estimate identity 1, every pair's match count is bounded by the shorter list, a sample of blocks equals the oracle."""
    from smoothxg_b200 import adaptive, synth
    batch = synth.make_batch(n_blocks=256, n_seqs=32, length=2000, divergence=0.02, seed=5)
    fb = adaptive.from_codes(batch)
    res = adaptive.block_identity(fb, kmer=17, want_pairs=True)
    assert np.all(res["n_kept"] == 32)
    assert np.all((res["threshold"] >= np.float32(0.7)) & (res["threshold"] <= np.float32(1.0)))
    assert res["stats"]["n_pairs"] == 256 * 496 and res["stats"]["kernel_launches"] == 3 * res["stats"]["n_chunks"]
    lens = fb.seq_len.reshape(256, 32) - 17
    iu = np.triu_indices(32, 1)
    bound = np.minimum(lens[:, iu[0]], lens[:, iu[1]]).reshape(-1)
    assert np.all(res["pair_common"] <= bound)
    for b in (0, 17, 255):
        ko, to, io, co = mash_oracle.block(fb.strings(b), 17)
        lo, hi = int(res["pair_off"][b]), int(res["pair_off"][b + 1])
        assert res["threshold"][b] == to and np.array_equal(res["pair_common"][lo:hi].astype(np.uint64), co) and np.array_equal(res["pair_identity"][lo:hi], io)
    twin = adaptive.flatten([[fb.strings(0)[0], fb.strings(0)[0]]])
    r2 = adaptive.block_identity(twin, kmer=17, want_pairs=True)
    assert r2["pair_identity"][0] == np.float32(1.0) and r2["threshold"][0] == np.float32(1.0)
