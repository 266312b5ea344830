"""the block graph smooth_abpoa returns -- build_odgi_abPOA (reference src/smooth.cpp:2442-2574),
odgi unchop, topological order, compact ids, path-supported edges (:545-620) -- from poa_b200_block_final_graph(), against what
the REFERENCE ITSELF returned for the 17 real DRB1 blocks in both alignment modes (harvested with integration/harvest.patch,
tests/golden/make_real_golden.py).  The reference's node ids after unchop depend on hash-map iteration order, so equality is
asserted up to renumbering: identical path walks (names, orientation, step count), a one-to-one node map along them, identical
node sequences and an identical edge set under that map.  CPU: POA results from the emulated device code; -m gpu: from the GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle.oracle import _Checker
from smoothxg_b200 import engine
from smoothxg_b200.shard import merge_parts
from tests.golden_io import engine_params, load_real_cases, load_real_meta, pd_params

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu_poa.cpp")
OUT = os.path.join(HERE, "emu", "_build", "libpoa_emu.so")
REAL = {c[0]: c for c in load_real_cases()}


def parse_gfa(text):
    nodes, edges, paths, order = {}, set(), {}, []
    for line in text.splitlines():
        f = line.split("\t")
        if f[0] == "S":
            nodes[int(f[1])] = f[2]
        elif f[0] == "L":
            a, oa, b, ob = int(f[1]), f[2], int(f[3]), f[4]
            if oa == "-" and ob == "-":
                a, b, oa, ob = b, a, "+", "+"
            assert oa == "+" and ob == "+", line
            edges.add((a, b))
        elif f[0] == "P":
            paths[f[1]] = [(int(s[:-1]), s[-1]) for s in f[2].split(",")] if f[2] not in ("", "*") else []
        elif f[0] == "#order":
            order = f[1:]
    return nodes, edges, paths, order


def check_block(fg, meta, gfa_text, label):
    """fg: engine.FinalGraph of our side; meta: (block id, padding, [(weight, revs, names)], reference GFA)."""
    nodes, edges, paths, order = parse_gfa(gfa_text)
    assert len(fg.node_seq) == len(nodes), f"{label}: {len(fg.node_seq)} nodes vs {len(nodes)} in the reference"
    ref2ours = {}
    seqs = meta[2]
    n_named = 0
    for i, (w, revs, names) in enumerate(seqs):
        ours = fg.path(i).tolist()
        for rev, name in zip(revs, names):
            want = paths[name]
            walk = ours[::-1] if rev else ours
            assert len(want) == len(walk), f"{label}: path {name}: {len(walk)} steps vs {len(want)}"
            for (rid, ro), oid in zip(want, walk):
                assert ro == ("-" if rev else "+"), f"{label}: path {name} orientation"
                assert ref2ours.setdefault(rid, oid) == oid, f"{label}: path {name}: node map is not a function"
            n_named += 1
    cons = [n for n in order if n not in {nm for _, _, names in seqs for nm in names}]
    assert len(cons) == 1 and len(order) == n_named + 1
    want, walk = paths[cons[0]], fg.path(len(seqs)).tolist()
    assert len(want) == len(walk), f"{label}: consensus path: {len(walk)} steps vs {len(want)}"
    for (rid, ro), oid in zip(want, walk):
        assert ro == "+" and ref2ours.setdefault(rid, oid) == oid, f"{label}: consensus path"
    assert len(ref2ours) == len(nodes) and len(set(ref2ours.values())) == len(nodes), f"{label}: node map is not one-to-one"
    for rid, oid in ref2ours.items():
        assert nodes[rid] == fg.node_seq[oid - 1], f"{label}: node {rid} sequence"
    assert {(ref2ours[a], ref2ours[b]) for a, b in edges} == set(zip(fg.edge_from.tolist(), fg.edge_to.tolist())), f"{label}: edge sets differ"
    # ids are a topological order: every edge goes up
    assert all(a < b for a, b in zip(fg.edge_from.tolist(), fg.edge_to.tolist())), f"{label}: ids are not topologically ordered"


@pytest.fixture(scope="module")
def wire():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    csrc = os.path.join(HERE, "..", "smoothxg_b200", "csrc")
    deps = [SRC] + [os.path.join(csrc, f) for f in ("poa_core.cuh", "poa_fill16.cuh", "poa_host.hpp", "poa_wire.hpp")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O1", "-fPIC", "-shared", "-std=c++17", "-I/usr/local/cuda/include", "-o", OUT, SRC])
    return _Checker(C.CDLL(OUT), "emu_poa_block_wire", "emu_free")


@pytest.mark.parametrize("mode,blocks", [("drb1_global", (0, 1, 5, 9, 13, 16)), ("drb1_local", (2, 8, 14))])
def test_final_graph_matches_reference_output(wire, mode, blocks):
    _, batch, p, _ = REAL[mode]
    meta = load_real_meta(mode)
    for b in blocks:
        w = wire.poa_block(pd_params(p), *batch.block(b)).raw
        hdr, arena = merge_parts(1, [(np.array([0]), w[:engine.HDR_WORDS], w[engine.HDR_WORDS:])])
        res = engine.result_from_parts(hdr, arena)
        check_block(res.final_graph(0, meta[b][1], True), meta[b], meta[b][3], f"{mode} block {meta[b][0]}")
        res.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["drb1_global", "drb1_local"])
def test_final_graph_matches_reference_output_gpu(mode):
    _, batch, p, _ = REAL[mode]
    meta = load_real_meta(mode)
    eng = engine.PoaEngine(device=0)
    res = eng.run_batch(batch, engine_params(p))
    for b in range(batch.n_blocks):
        check_block(res.final_graph(b, meta[b][1], True), meta[b], meta[b][3], f"{mode} block {meta[b][0]}")
    res.close(); eng.close()
