"""CPU: poa_b200_block_graph() -- the per-block graph smoothxg's build_odgi_abPOA leaves behind (reference
src/smooth.cpp:2442-2574) -- against a literal restatement of that function on a dict graph.
(The restatement is ours; what pins this stage to the reference itself is tests/test_final_graph.py, which compares the graph
after the next stage with the block graphs the reference binary returned for real blocks.)
The POA results come from the emulated device code (wire format -> poa_b200_result_from_parts), so the test needs
no GPU; the -m gpu suite repeats it on results produced by the CUDA path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle.oracle import _Checker
from smoothxg_b200 import engine
from smoothxg_b200.shard import merge_parts
from tests.golden_io import load_cases, pd_params

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu_poa.cpp")
OUT = os.path.join(HERE, "emu", "_build", "libpoa_emu.so")
CASES = {c[0]: c for c in load_cases()}


def build_odgi_restated(v, padding_len, include_consensus):
    """src/smooth.cpp:2442-2574 step by step: create every node and edge in a Kahn walk, embed read paths with the
    padding trimmed, embed the consensus over still-covered nodes, drop edges and nodes no path uses."""
    n = v.n_node
    if n <= 2:
        return [], [], [[] for _ in range(v.n_seq + (1 if include_consensus else 0))]
    in_off = np.concatenate([[0], np.cumsum(v.in_n)]); out_off = np.concatenate([[0], np.cumsum(v.out_n)])
    nodes, edges = [], []                      # creation order
    indeg = v.in_n.copy(); q = [0]; qh = 0
    while qh < len(q):                          # :2463-2511
        cur = q[qh]; qh += 1
        if cur == 1:
            break
        if cur != 0:
            nodes.append(cur - 1)
            for k in range(in_off[cur], in_off[cur + 1]):
                pre = int(v.in_id[k])
                if pre != 0:
                    edges.append((pre - 1, cur - 1))
        for k in range(out_off[cur], out_off[cur + 1]):
            o = int(v.out_id[k]); indeg[o] -= 1
            if indeg[o] == 0:
                q.append(o)
    paths, off = [], 0
    steps = {}                                  # node -> number of path steps
    for i in range(v.n_seq):                    # :2513-2532
        ln = int(v.path_len[i])
        p = [int(x) - 1 for x in v.path_node[off + padding_len: off + max(ln - padding_len, padding_len)]] if ln - padding_len > padding_len else []
        for x in p:
            steps[x] = steps.get(x, 0) + 1
        paths.append(p); off += ln
    if include_consensus:                       # :2534-2549
        p = [int(c) - 1 for c in v.cons_node if steps.get(int(c) - 1, 0) > 0]
        for x in p:
            steps[x] = steps.get(x, 0) + 1
        paths.append(p)
    # :2559-2565 removes nothing: odgi's find_edges_exceeding_depth_limits(min_depth=1) only inspects edges some path walks
    # (deps/odgi/src/algorithms/depth.cpp:17-51); edges go away only with an uncovered node (:2567-2573, destroy_handle)
    edges = [e for e in edges if steps.get(e[0], 0) > 0 and steps.get(e[1], 0) > 0]
    nodes = [x for x in nodes if steps.get(x, 0) > 0]  # :2567-2573
    return nodes, edges, paths


@pytest.fixture(scope="module")
def wire():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    csrc = os.path.join(HERE, "..", "smoothxg_b200", "csrc")
    deps = [SRC] + [os.path.join(csrc, f) for f in ("poa_core.cuh", "poa_fill16.cuh", "poa_host.hpp", "poa_wire.hpp")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O1", "-fPIC", "-shared", "-std=c++17", "-I/usr/local/cuda/include", "-o", OUT, SRC])
    return _Checker(C.CDLL(OUT), "emu_poa_block_wire", "emu_free")


@pytest.mark.parametrize("name", ["abpoa_seq_fa_global", "edge_shapes", "syn_global_band_msa_w", "syn_indel", "syn_local", "affine_global_band"])
@pytest.mark.parametrize("padding,cons", [(0, True), (7, True), (40, False), (5000, True)])
def test_block_graph_matches_build_odgi(wire, name, padding, cons):
    _, batch, p, _ = CASES[name]
    parts = []
    for b in range(batch.n_blocks):
        w = wire.poa_block(pd_params(p), *batch.block(b)).raw
        parts.append((np.array([b]), w[:engine.HDR_WORDS], w[engine.HDR_WORDS:]))
    hdr, arena = merge_parts(batch.n_blocks, parts)
    res = engine.result_from_parts(hdr, arena)
    for b in range(batch.n_blocks):
        v = res.block(b)
        use_cons = cons and v.cons_len >= 0
        g = res.block_graph(b, padding, use_cons)
        nodes, edges, paths = build_odgi_restated(v, padding, use_cons)
        assert g.node_id.tolist() == nodes
        assert list(zip(g.edge_from.tolist(), g.edge_to.tolist())) == edges
        assert len(g.path_off) - 1 == len(paths)
        for i, pth in enumerate(paths):
            assert g.path(i).tolist() == pth
        want_bases = bytes("ACGTN"[int(v.base[x + 1])].encode()[0] for x in nodes)
        assert g.node_base == want_bases
        # every read still spells its (trimmed) sequence through the kept nodes
        lens, bases, _ = batch.block(b)
        off = 0
        for i in range(v.n_seq):
            ln = int(lens[i])
            if int(v.path_len[i]) == ln and ln - 2 * padding > 0:
                seq = bases[off + padding: off + ln - padding]
                assert np.array_equal(v.base[g.path(i) + 1].astype(np.uint8), seq)
            off += ln
    res.close()


@pytest.mark.parametrize("name", ["abpoa_seq_fa_global", "edge_shapes", "syn_indel", "syn_local"])
def test_block_hash_matches_unmodified_abpoa(wire, name):
    """poa_b200_result_block_hash() hashes the same byte stream as oracle/ref_shim.c:ref_poa_batch_timed does from
    abpoa_graph_t (node_n, base, out ids, out weights): bench.py compares its CPU sample with the GPU result this way."""
    from oracle.oracle import RefAbpoa, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    _, batch, p, _ = CASES[name]
    parts = []
    for b in range(batch.n_blocks):
        w = wire.poa_block(pd_params(p), *batch.block(b)).raw
        parts.append((np.array([b]), w[:engine.HDR_WORDS], w[engine.HDR_WORDS:]))
    hdr, arena = merge_parts(batch.n_blocks, parts)
    res = engine.result_from_parts(hdr, arena)
    _, want = RefAbpoa().batch_timed(pd_params(p), batch, n_threads=2, want_hash=True)
    got = np.array([res.block_hash(b) for b in range(batch.n_blocks)], dtype=np.uint64)
    assert np.array_equal(got, want)
    res.close()
