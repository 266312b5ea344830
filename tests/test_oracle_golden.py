"""CPU: the scalar restatement (oracle/poa_oracle.c) against the committed golden vectors produced by
the unmodified vendored abPOA -- graph, read paths, consensus, MSA, scores, cigars, band cell counts."""
import numpy as np
import pytest

from tests.golden_io import load_cases, load_real_cases, pd_params

CASES = load_cases()
REAL = load_real_cases()


@pytest.mark.parametrize("name,batch,p,dumps", REAL, ids=[c[0] for c in REAL])
def test_oracle_matches_real_blocks(oracle, name, batch, p, dumps):
    """Real DRB1 blocks as smoothxg hands them to abPOA (N padding, dedup weights, long predecessor edges, in-degree up to 6)."""
    assert batch.n_blocks == 17 and int(batch.weight.max()) > 1 and int((batch.bases == 4).sum()) > 0
    for b in range(batch.n_blocks):
        got = oracle.poa_block(pd_params(p), *batch.block(b))
        assert got is not None
        assert np.array_equal(got.raw, dumps[b].raw), f"{name} block {b}"


@pytest.mark.parametrize("name,batch,p,dumps", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_golden(oracle, name, batch, p, dumps):
    for b in range(batch.n_blocks):
        got = oracle.poa_block(pd_params(p), *batch.block(b))
        assert got is not None
        assert np.array_equal(got.raw, dumps[b].raw), f"{name} block {b}"


def test_oracle_matches_full_size_local_block(oracle):
    """One configs[2]-shaped block (32 x 2 kb) in LOCAL mode -- what plain -A gives (src/main.cpp:487) -- against the
    unmodified abPOA's per-section digests (tests/golden/make_deep_golden.py)."""
    from oracle.oracle import make_params
    from smoothxg_b200.synth import make_batch
    from tests.golden_io import DEEP_CASES, deep_mismatches
    kw, pk = DEEP_CASES["local_32x2kb"]
    got = oracle.poa_block(make_params(**pk), *make_batch(**kw).block(0))
    assert deep_mismatches("local_32x2kb", got) == []


def test_oracle_lane_count_only_moves_junk(oracle):
    """The reference's SIMD lane count only shifts the band start over -inf cells: results equal."""
    name, batch, p, dumps = next(c for c in CASES if c[0] == "syn_indel")
    try:
        for pn16, pn32 in ((8, 4), (16, 8)):
            oracle.set_lane_counts(pn16, pn32)
            for b in range(batch.n_blocks):
                got = oracle.poa_block(pd_params(p), *batch.block(b))
                assert np.array_equal(got.result_part(), dumps[b].result_part())
    finally:
        oracle.set_lane_counts(32, 16)


def test_gap_mode_follows_abpoa_set_gap_mode(oracle):
    """gap_open1 == 0 -> linear, gap_open2 == 0 -> affine, else convex (abpoa_align.c:87-91): the three kernels give
    different graphs on an indel-rich block, and each is pinned by its own golden case above."""
    from oracle.oracle import make_params
    name, batch, _, _ = next(c for c in CASES if c[0] == "syn_indel")
    lens = [oracle.poa_block(make_params(**kw), *batch.block(0)).raw.tobytes()
            for kw in (dict(), dict(gap_open2=0, gap_ext2=0), dict(gap_open1=0, gap_ext1=2, gap_open2=0, gap_ext2=0))]
    assert len(set(lens)) == 3
