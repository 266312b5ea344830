"""CPU, only where oracle/_ref exists (the build container, or a GPU box that received the prebuilt .so):
the restatement against the live unmodified abPOA on fresh seeded inputs, and SSE4.1 = AVX2 = AVX-512."""
import numpy as np
import pytest

from oracle.oracle import RefAbpoa, make_params, ref_available, _cpu_has
from smoothxg_b200.synth import make_batch

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("kw,pk", [
    (dict(n_blocks=2, n_seqs=10, length=700, seed=101), dict()),
    (dict(n_blocks=2, n_seqs=6, length=400, seed=102), dict(local=True, out_msa=True)),
    (dict(n_blocks=2, n_seqs=8, length=900, seed=103, indel_prob=0.7, dup_weights=True, n_frac=0.02), dict(out_msa=True)),
    (dict(n_blocks=2, n_seqs=6, length=300, seed=104, divergence=0.15), dict(banded=False)),
    # affine (four-value -p in smoothxg) and linear gap kernels
    (dict(n_blocks=2, n_seqs=10, length=700, seed=111, indel_prob=0.4, indel_len=(30, 200)), dict(gap_open2=0, gap_ext2=0)),
    (dict(n_blocks=2, n_seqs=6, length=400, seed=112, n_frac=0.02, dup_weights=True), dict(gap_open2=0, gap_ext2=0, local=True, out_msa=True)),
    (dict(n_blocks=2, n_seqs=6, length=300, seed=113, divergence=0.15), dict(gap_open1=9, gap_ext1=1, gap_open2=0, gap_ext2=0, mismatch=3, banded=False)),
    (dict(n_blocks=2, n_seqs=10, length=700, seed=114, indel_prob=0.4, indel_len=(30, 200)), dict(gap_open1=0, gap_ext1=2, gap_open2=0, gap_ext2=0)),
    (dict(n_blocks=2, n_seqs=6, length=400, seed=115), dict(gap_open1=0, gap_ext1=3, gap_open2=0, gap_ext2=0, local=True, out_msa=True)),
    (dict(n_blocks=2, n_seqs=6, length=300, seed=116, divergence=0.15), dict(gap_open1=0, gap_ext1=2, gap_open2=0, gap_ext2=0, banded=False)),
])
def test_restatement_equals_reference(oracle, kw, pk):
    ref = RefAbpoa()
    batch = make_batch(**kw)
    p = make_params(**pk)
    for b in range(batch.n_blocks):
        a, r = oracle.poa_block(p, *batch.block(b)), ref.poa_block(p, *batch.block(b))
        assert np.array_equal(a.raw, r.raw)


def test_uninstrumented_call_gives_same_result(oracle):
    """instrument=0 calls abpoa_poa() itself, the exact call smoothxg makes (src/smooth.cpp:337)."""
    ref = RefAbpoa()
    batch = make_batch(n_blocks=2, n_seqs=8, length=500, seed=105)
    p = make_params(out_msa=True)
    for b in range(batch.n_blocks):
        a = ref.poa_block(p, *batch.block(b), instrument=False)
        c = ref.poa_block(p, *batch.block(b), instrument=True)
        assert np.array_equal(a.result_part(), c.result_part())


def test_cross_isa_identity():
    isas = ["sse41"] + (["avx2"] if _cpu_has("avx2") else []) + (["avx512"] if _cpu_has("avx512bw") else [])
    isas = [i for i in isas if ref_available(i)]
    if len(isas) < 2:
        pytest.skip("fewer than two ISA builds runnable here")
    batch = make_batch(n_blocks=2, n_seqs=8, length=600, seed=106, indel_prob=0.5)
    p = make_params()
    base = [RefAbpoa(isas[0]).poa_block(p, *batch.block(b)) for b in range(batch.n_blocks)]
    for isa in isas[1:]:
        r = RefAbpoa(isa)
        for b in range(batch.n_blocks):
            assert np.array_equal(r.poa_block(p, *batch.block(b)).result_part(), base[b].result_part())


def test_batch_timed_runs():
    ref = RefAbpoa()
    batch = make_batch(n_blocks=4, n_seqs=4, length=200, seed=107)
    secs, h = ref.batch_timed(make_params(), batch, n_threads=2, want_hash=True)
    assert secs > 0 and len(set(h.tolist())) == 4
