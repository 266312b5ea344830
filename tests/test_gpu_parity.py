"""GPU (-m gpu): the CUDA path through the C ABI against (a) the committed golden vectors of the unmodified
abPOA, (b) the oracle on fresh seeded inputs, (c) size-independent properties at BASELINE.json's full sizes.
Bit-exact everywhere: integer DP, node ids, edge order, weights, paths, consensus, MSA, scores, cigars and the
in-band cell count."""
import numpy as np
import pytest

from oracle.oracle import make_params as oracle_params
from smoothxg_b200 import engine as E
from smoothxg_b200 import shard, synth
from tests.golden_io import engine_params, load_cases, load_real_cases, pd_params
from tests.helpers import first_diff, view_to_dump

pytestmark = pytest.mark.gpu
CASES = load_cases()


def _check_batch(eng, batch, eparams, want_dumps, label):
    res = eng.run_batch(batch, eparams)
    for b in range(batch.n_blocks):
        got = view_to_dump(res.block(b))
        assert np.array_equal(got.compare_part(), want_dumps[b].compare_part()), f"{label} block {b}: {first_diff(want_dumps[b], got)}"
    st = res.stats()
    res.close()
    return st


@pytest.mark.parametrize("warps", [1, 2, 4, 8])
def test_golden_vectors(warps):
    eng = E.PoaEngine(device=0, emit_cigar=True, warps_per_block=warps)
    for name, batch, p, dumps in CASES:
        _check_batch(eng, batch, engine_params(p), dumps, f"{name}/w{warps}")
    eng.close()


@pytest.mark.parametrize("warps", [1, 2, 4, 8])
def test_real_drb1_blocks(warps):
    """BASELINE configs[0]'s data on the abPOA path: the 17 blocks smoothxg forms from DRB1-3123 (-l 1100), exactly as handed
    to abPOA -- global-banded (-A -Z) and local (-A) -- against the unmodified abPOA's dumps."""
    eng = E.PoaEngine(device=0, emit_cigar=True, warps_per_block=warps)
    for name, batch, p, dumps in load_real_cases():
        st = _check_batch(eng, batch, engine_params(p), dumps, f"{name}/w{warps}")
        assert st["inband_cells"] == sum(d.inband_cells for d in dumps)
    eng.close()


@pytest.mark.parametrize("name", ["local_32x2kb", "deep_256x8kb"])
def test_full_size_single_blocks(name):
    """One full-size configs[3] block (256 x 8 kb: int16 -> int32 switch mid-block, 28 592 rows) and one 32 x 2 kb block in
    local mode, against the unmodified abPOA's per-section digests (graph, edge order, weights, paths, scores, cigars)."""
    from tests.golden_io import DEEP_CASES, deep_mismatches
    kw, pk = DEEP_CASES[name]
    batch = synth.make_batch(**kw)
    eng = E.PoaEngine(device=0, emit_cigar=True)
    res = eng.run_batch(batch, E.make_params(**pk))
    assert deep_mismatches(name, view_to_dump(res.block(0))) == []
    res.close(); eng.close()


@pytest.mark.parametrize("kw,pk", [
    (dict(n_blocks=24, n_seqs=16, length=1000, seed=201), dict()),
    (dict(n_blocks=8, n_seqs=12, length=1500, seed=202, indel_prob=0.6, indel_len=(100, 700), dup_weights=True, n_frac=0.01), dict(out_msa=True)),
    (dict(n_blocks=8, n_seqs=8, length=700, seed=203), dict(local=True, out_msa=True)),
    (dict(n_blocks=8, n_seqs=8, length=500, seed=204, divergence=0.12), dict(banded=False)),
    (dict(n_blocks=6, n_seqs=32, length=2000, seed=205), dict()),
    (dict(n_blocks=4, n_seqs=10, length=800, seed=206, divergence=0.001), dict(out_msa=True)),
    # rows wider than the packed fill's shared-memory ring (6 x 256 columns): unbanded and local, int16
    (dict(n_blocks=3, n_seqs=5, length=2100, seed=207), dict(banded=False)),
    (dict(n_blocks=3, n_seqs=5, length=1800, seed=208, indel_prob=0.3, indel_len=(50, 300)), dict(local=True, out_msa=True)),
    # long band with large indels: predecessor rows far back, bands that jump
    (dict(n_blocks=3, n_seqs=6, length=3000, seed=209, indel_prob=0.5, indel_len=(100, 600)), dict()),
])
def test_fresh_inputs_vs_oracle(engine, oracle, kw, pk):
    batch = synth.make_batch(**kw)
    want = oracle.poa_batch(oracle_params(**pk), batch)
    st = _check_batch(engine, batch, E.make_params(**pk), want, str(kw))
    assert st["inband_cells"] == sum(d.inband_cells for d in want)
    assert st["edge_row_cells"] == sum(d.edge_rows for d in want)


@pytest.mark.parametrize("isa_bits", [1, 2])
def test_lane_count_option_gives_identical_results(isa_bits):
    """poa_b200_engine_opts_t::flags bits 4-5 select the SIMD width whose band-start rounding is reproduced (AVX2 = 16 int16 lanes,
    SSE4.1 / NEON = 8; default AVX-512BW = 32).  The rule only moves a row's band start over -inf cells, so the results are the
    same as the golden vectors pinned with the AVX-512 build (SURVEY 6.2 found the three abPOA builds identical too) -- the
    in-band cell COUNT may differ, results may not."""
    eng = E.PoaEngine(device=0, emit_cigar=True, flags=isa_bits << 4)
    for name, batch, p, dumps in load_real_cases() + [c for c in CASES if c[0] in ("syn_indel", "syn_global_band", "abpoa_heter_fa_global")]:
        res = eng.run_batch(batch, engine_params(p))
        for b in range(batch.n_blocks):
            assert np.array_equal(view_to_dump(res.block(b)).result_part(), dumps[b].result_part()), f"{name} block {b}"
        res.close()
    eng.close()


def test_engine_trim_returns_pooled_memory():
    import torch
    eng = E.PoaEngine(device=0)
    batch = synth.make_batch(n_blocks=64, n_seqs=8, length=500, seed=5)
    eng.run_batch(batch, E.make_params()).close()
    free0 = torch.cuda.mem_get_info(0)[0]
    eng.trim()
    assert torch.cuda.mem_get_info(0)[0] > free0  # workspace / arena / input buffers went back to the driver
    eng.run_batch(batch, E.make_params()).close()  # and the engine still works
    eng.close()


def test_wide_wire_format_and_arena_retry(engine, oracle):
    """Dedup weights whose sum passes 65 535 force the 32-bit form of the result body (WIRE_WIDE): twice the words of the
    first-guess arena estimate for unrelated sequences, so some blocks overflow the arena and are re-run with the exact size the
    kernel reported.  Results must not depend on any of that."""
    batch = synth.make_batch(n_blocks=12, n_seqs=24, length=260, seed=77, divergence=0.7)
    batch.weight[:] = 3000 + (np.arange(batch.weight.shape[0]) % 7) * 500
    pk = dict(out_msa=True)
    want = oracle.poa_batch(oracle_params(**pk), batch)
    st = _check_batch(engine, batch, E.make_params(**pk), want, "wide")
    assert st["inband_cells"] == sum(d.inband_cells for d in want)


def test_int32_scores_and_mid_block_switch(engine, oracle):
    """max(qlen, rows) > 16361 switches abPOA to 32-bit scores (abpoa_align_simd.c:1293-1302), also mid-block."""
    for kw in (dict(n_blocks=1, n_seqs=3, length=17000, seed=21), dict(n_blocks=1, n_seqs=6, length=8000, seed=22, divergence=0.3)):
        batch = synth.make_batch(**kw)
        want = oracle.poa_batch(oracle_params(), batch)
        assert want[0].n_node > 16400
        _check_batch(engine, batch, E.make_params(), want, str(kw))


@pytest.mark.parametrize("warps", [0, 1])
def test_deep_block(oracle, warps):
    """BASELINE.json configs[3] in miniature (deep-block stress): 96 sequences per block, the graph grows to ~3x the
    sequence length, the first-guess workspace overflows and the block is re-run with a larger one.  warps=0 lets the
    engine choose (several warps per block for a two-block batch: generic fill), warps=1 forces the packed fill."""
    batch = synth.make_batch(n_blocks=2, n_seqs=96, length=2500, seed=240, divergence=0.03)
    want = oracle.poa_batch(oracle_params(), batch)
    assert want[0].n_node > 2 * 2500
    eng = E.PoaEngine(device=0, emit_cigar=True, warps_per_block=warps)
    st = _check_batch(eng, batch, E.make_params(), want, f"deep/w{warps}")
    assert st["inband_cells"] == sum(d.inband_cells for d in want)
    eng.close()


def test_workspace_retry_gives_identical_results(oracle):
    """Blocks that exhaust the first-guess workspace are re-run with larger ones; results must not change."""
    batch = synth.make_batch(n_blocks=6, n_seqs=8, length=600, seed=9, divergence=0.25)
    want = oracle.poa_batch(oracle_params(), batch)
    eng = E.PoaEngine(device=0, emit_cigar=True, slab_rows_factor=1.0)
    st = _check_batch(eng, batch, E.make_params(), want, "retry")
    assert st["retried_blocks"] > 0 and st["kernel_launches"] >= 2
    eng.close()


def test_poa_block_call_shape(engine, oracle):
    """poa_b200_poa_block takes abpoa_poa's argument shapes (seqs[], seq_lens[], weights[])."""
    batch = synth.make_batch(n_blocks=1, n_seqs=7, length=300, seed=31, dup_weights=True)
    lens, bases, wts = batch.block(0)
    res = engine.poa_block(batch.block_seqs(0), wts, E.make_params(out_msa=True))
    want = oracle.poa_block(oracle_params(out_msa=True), lens, bases, wts)
    assert np.array_equal(view_to_dump(res.block(0)).compare_part(), want.compare_part())


@pytest.mark.parametrize("pk", [dict(gap_open2=0, gap_ext2=0), dict(gap_open1=0, gap_ext1=2, gap_open2=0, gap_ext2=0)], ids=["affine", "linear"])
@pytest.mark.parametrize("mode", [dict(), dict(local=True, out_msa=True), dict(banded=False)], ids=["band", "local", "unbanded"])
@pytest.mark.parametrize("warps", [1, 4])
def test_affine_and_linear_gap_modes(oracle, pk, mode, warps):
    """abpoa_set_gap_mode (abpoa_align.c:87-91): four-value -p in smoothxg gives the affine kernel; gap_open1 = 0 the linear one."""
    batch = synth.make_batch(n_blocks=6, n_seqs=8, length=600, seed=220, indel_prob=0.3, indel_len=(20, 150), n_frac=0.01, dup_weights=True)
    kw = dict(pk, **mode)
    want = oracle.poa_batch(oracle_params(**kw), batch)
    eng = E.PoaEngine(device=0, emit_cigar=True, warps_per_block=warps)
    st = _check_batch(eng, batch, E.make_params(**kw), want, str(kw))
    assert st["inband_cells"] == sum(d.inband_cells for d in want)
    eng.close()


def test_zero_gap_extension_is_refused(engine):
    """e1 = e2 = 0 leaves the reference's 16-bit scores without head room (inf_min = INT16_MIN + ..., abpoa_align_simd.c:1295)."""
    batch = synth.make_batch(n_blocks=1, n_seqs=2, length=50, seed=1)
    with pytest.raises(E.PoaError) as e:
        engine.run_batch(batch, E.make_params(gap_ext1=0, gap_ext2=0))
    assert e.value.code == E.EUNSUP


def test_block_graph_from_gpu_results(engine):
    """poa_b200_block_graph (build_odgi_abPOA equivalent, src/smooth.cpp:2442-2574) on results of the CUDA path."""
    from tests.test_block_graph import build_odgi_restated
    batch = synth.make_batch(n_blocks=5, n_seqs=9, length=500, seed=230, indel_prob=0.3, dup_weights=True)
    res = engine.run_batch(batch, E.make_params(out_msa=True))
    for b in range(batch.n_blocks):
        v = res.block(b)
        for padding in (0, 31):
            g = res.block_graph(b, padding, True)
            nodes, edges, paths = build_odgi_restated(v, padding, True)
            assert g.node_id.tolist() == nodes and list(zip(g.edge_from.tolist(), g.edge_to.tolist())) == edges
            assert [g.path(i).tolist() for i in range(len(paths))] == paths
    res.close()


def test_empty_batch(engine):
    batch = synth.PoaBatch.from_blocks([])
    res = engine.run_batch(batch, E.make_params())
    assert len(res) == 0


def _properties(batch, res, blocks):
    """Size-independent invariants of a POA result (hold for any input): every read's path spells the read;
    the source's out-weights and the sink's in-weights sum to the total read weight; every edge list is sorted by
    weight (abpoa_graph.c:192-219); in- and out-lists describe the same edges; the consensus is a source-to-sink walk."""
    for b in blocks:
        v = res.block(b)
        assert v.status == 0
        lens, bases, wts = batch.block(b)
        assert np.array_equal(v.path_len, lens)
        assert np.array_equal(v.base[v.path_node].astype(np.uint8), bases)
        out_off = np.concatenate([[0], np.cumsum(v.out_n)]); in_off = np.concatenate([[0], np.cumsum(v.in_n)])
        assert v.out_w[out_off[0]:out_off[1]].sum() == wts.sum() == v.in_w[in_off[1]:in_off[2]].sum()
        src = np.repeat(np.arange(v.n_node), v.out_n); dst = np.repeat(np.arange(v.n_node), v.in_n)
        e_out = np.stack([src, v.out_id, v.out_w], 1); e_in = np.stack([v.in_id, dst, v.in_w], 1)
        assert np.array_equal(e_out[np.lexsort(e_out.T[::-1])], e_in[np.lexsort(e_in.T[::-1])])
        for off, w in ((out_off, v.out_w), (in_off, v.in_w)):
            d = np.diff(w); brk = np.zeros(d.shape[0], bool); idx = off[1:-1] - 1
            brk[idx[(idx >= 0) & (idx < d.shape[0])]] = True
            assert np.all((d <= 0) | brk)
        # total path weight through every node = its in-weight = its out-weight
        win = np.add.reduceat(np.concatenate([v.in_w, [0]]), np.minimum(in_off[:-1], v.in_w.shape[0]))[2:]
        wout = np.add.reduceat(np.concatenate([v.out_w, [0]]), np.minimum(out_off[:-1], v.out_w.shape[0]))[2:]
        assert np.array_equal(win, wout)
        if v.cons_len > 0:
            c = v.cons_node
            eset = set(zip(src.tolist(), v.out_id.tolist()))
            assert (0, int(c[0])) in eset and (int(c[-1]), 1) in eset
            assert all((int(a), int(b_)) in eset for a, b_ in zip(c[:-1], c[1:]))


def test_packed_fill_equals_generic_fill(oracle):
    """engine flag bit 0 turns the packed 16-bit fill off; both code paths must give the oracle's result."""
    batch = synth.make_batch(n_blocks=12, n_seqs=10, length=900, seed=210, indel_prob=0.2)
    want = oracle.poa_batch(oracle_params(out_msa=True), batch)
    for flags in (0, 1):
        eng = E.PoaEngine(device=0, emit_cigar=True, flags=flags, warps_per_block=1)
        _check_batch(eng, batch, E.make_params(out_msa=True), want, f"flags={flags}")
        eng.close()


def test_full_size_config1_properties_and_determinism():
    """BASELINE.json configs[1]: 1000 blocks x 16 seqs x 1 kb, int16.  Properties on every block; the result is
    identical for 1 and 4 warps per block (scheduling-independent) -- checksum of checksums."""
    batch = synth.make_batch(**synth.CONFIGS["config1_1000x16x1k"], seed=1000)
    sums = []
    for warps in (1, 4):
        eng = E.PoaEngine(device=0, warps_per_block=warps)
        res = eng.run_batch(batch, E.make_params())
        if warps == 1:
            _properties(batch, res, range(batch.n_blocks))
        h = 0
        for b in range(batch.n_blocks):
            v = res.block(b)
            h = hash((h, v.n_node, v.out_id.tobytes(), v.out_w.tobytes(), v.path_node.tobytes(), v.cons_node.tobytes(), v.inband_cells))
        sums.append(h)
        res.close(); eng.close()
    assert sums[0] == sums[1]


def test_full_size_config2_properties(oracle):
    """BASELINE.json configs[2]: 10 000 blocks x 32 seqs x 2 kb, adaptive band: all blocks finish, invariants hold on a
    sample, and a few blocks are compared with the oracle outright."""
    batch = synth.make_batch(**synth.CONFIGS["config2_10000x32x2k"], seed=1000)
    eng = E.PoaEngine(device=0)
    res = eng.run_batch(batch, E.make_params())
    st = res.stats()
    assert st["retried_blocks"] == 0 or st["kernel_launches"] > 1
    sample = list(range(0, batch.n_blocks, 97))
    _properties(batch, res, sample)
    for b in range(0, batch.n_blocks, 10):
        assert res.block(b).status == 0
    for b in (0, 4999, 9999):
        want = oracle.poa_block(oracle_params(), *batch.block(b), instrument=False)
        assert np.array_equal(view_to_dump(res.block(b)).result_part(), want.result_part())
    res.close(); eng.close()


def test_run_sharded_single_rank_equals_run_batch(engine):
    batch = synth.make_batch(n_blocks=10, n_seqs=6, length=300, seed=41)
    p = E.make_params(out_msa=True)
    a = engine.run_batch(batch, p)
    b = shard.run_sharded(engine, batch, p)
    for i in range(batch.n_blocks):
        assert np.array_equal(view_to_dump(a.block(i)).compare_part(), view_to_dump(b.block(i)).compare_part())
