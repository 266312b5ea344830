#!/usr/bin/env python
"""Generate tests/golden/abpoa_golden.npz by running the UNMODIFIED vendored abPOA v1.5.4
(oracle/_ref, built from /root/reference/deps/abPOA by oracle/Makefile) the way smooth_abpoa drives it
(oracle/ref_shim.c).  Run in the build container (needs /root/reference); the npz is committed so the
GPU box, which has no /root/reference, can still pin the oracle and the CUDA path to the reference.

Upstream has no known-answer tests for abPOA, so these are our golden vectors:
  * abPOA's own test inputs: deps/abPOA/test_data/{seq,test,heter}.fa and example.c's second set
  * seeded synthetic blocks in every mode the hot path supports (convex / affine / linear gaps; global banded /
    unbanded / local, N bases, dedup weights, long indels, MSA on/off) and degenerate shapes (one sequence, empty block)
Each case stores the flat inputs, the parameter tuple and the canonical dump (oracle/poa_dump.h)
of an instrumented run: graph, read paths, consensus, MSA, per-sequence scores and cigars, band cells.
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import RefAbpoa, make_params  # noqa: E402
from smoothxg_b200.synth import PoaBatch, make_batch  # noqa: E402

REF = "/root/reference/deps/abPOA"


def read_fasta(path):
    seqs, cur = [], []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            if cur:
                seqs.append("".join(cur)); cur = []
        elif line:
            cur.append(line)
    if cur:
        seqs.append("".join(cur))
    return seqs


def example_c_seqs():
    txt = open(os.path.join(REF, "example.c")).read()
    body = txt[txt.index("\n    char seqs[10][100]"):]  # the live definition (the commented-out set is the same data as test_data/seq.fa)
    body = body[:body.index("};")]
    return [m for m in re.findall(r'^\s*"([ACGT]+)"', body, flags=re.M)]


def cases():
    P = dict  # parameter kwargs for make_params
    out = []
    seq_fa = read_fasta(os.path.join(REF, "test_data/seq.fa"))
    out.append(("abpoa_seq_fa_global", PoaBatch.from_strings([seq_fa]), P(out_msa=True)))
    out.append(("abpoa_seq_fa_local", PoaBatch.from_strings([seq_fa]), P(local=True, out_msa=True)))
    out.append(("abpoa_test_fa", PoaBatch.from_strings([read_fasta(os.path.join(REF, "test_data/test.fa"))]), P(out_msa=True)))
    heter = read_fasta(os.path.join(REF, "test_data/heter.fa"))
    out.append(("abpoa_heter_fa_global", PoaBatch.from_strings([heter]), P(out_msa=True)))
    out.append(("abpoa_heter_fa_unbanded", PoaBatch.from_strings([heter]), P(banded=False)))
    ex = example_c_seqs()
    assert len(ex) == 10
    out.append(("abpoa_example_c", PoaBatch.from_strings([ex]), P(out_msa=True)))
    out.append(("edge_shapes", PoaBatch.from_strings([["ACGTACGT", "ACGTTCGT", "ACGACGT"], ["A"], ["ACGT", "ACGT"], ["AC", "", "G"], [],
                                                      ["NNNNACGTNN", "ACGT", "NNNN"], ["ACGT" * 20, "TTTT" * 20]]), P(out_msa=True)))
    out.append(("edge_shapes_local", PoaBatch.from_strings([["ACGTACGT", "ACGTTCGT", "ACGACGT"], ["A"], ["ACGT" * 20, "TTTT" * 20, "GGGG"]]), P(local=True, out_msa=True)))
    out.append(("syn_global_band", make_batch(4, 8, 500, 0.02, seed=11), P()))
    out.append(("syn_global_band_msa_w", make_batch(3, 8, 400, 0.05, seed=12, n_frac=0.01, dup_weights=True), P(out_msa=True)))
    out.append(("syn_indel", make_batch(3, 8, 900, 0.02, seed=13, indel_prob=0.6, indel_len=(50, 400)), P()))
    out.append(("syn_local", make_batch(3, 6, 400, 0.03, seed=14), P(local=True, out_msa=True)))
    out.append(("syn_unbanded", make_batch(3, 6, 300, 0.10, seed=15), P(banded=False)))
    out.append(("syn_divergent", make_batch(2, 6, 400, 0.25, seed=16), P(out_msa=True)))
    out.append(("syn_presets", make_batch(2, 6, 400, 0.02, seed=17), P(match=1, mismatch=19, gap_open1=39, gap_ext1=3, gap_open2=81, gap_ext2=1)))
    out.append(("syn_presets2", make_batch(2, 6, 400, 0.08, seed=18), P(match=1, mismatch=7, gap_open1=11, gap_ext1=2, gap_open2=33, gap_ext2=1)))
    # affine gaps: what smoothxg passes when -p has four values (src/main.cpp:353-359); linear gaps: gap_open1 == 0
    AFF, LIN = dict(gap_open2=0, gap_ext2=0), dict(gap_open1=0, gap_ext1=2, gap_open2=0, gap_ext2=0)
    out.append(("affine_seq_fa", PoaBatch.from_strings([seq_fa]), P(out_msa=True, **AFF)))
    out.append(("affine_global_band", make_batch(3, 8, 400, 0.03, seed=21, indel_prob=0.3, indel_len=(20, 120)), P(**AFF)))
    out.append(("affine_local", make_batch(3, 6, 300, 0.04, seed=22, n_frac=0.01, dup_weights=True), P(local=True, out_msa=True, **AFF)))
    out.append(("affine_unbanded", make_batch(2, 6, 300, 0.10, seed=23), P(banded=False, match=2, mismatch=5, gap_open1=8, gap_ext1=1, gap_open2=0, gap_ext2=0)))
    out.append(("linear_seq_fa", PoaBatch.from_strings([seq_fa]), P(out_msa=True, **LIN)))
    out.append(("linear_global_band", make_batch(3, 8, 400, 0.03, seed=24, indel_prob=0.3, indel_len=(20, 120)), P(**LIN)))
    out.append(("linear_local", make_batch(3, 6, 300, 0.04, seed=25), P(local=True, out_msa=True, **LIN)))
    out.append(("linear_unbanded", make_batch(2, 6, 300, 0.10, seed=26), P(banded=False, gap_open1=0, gap_ext1=3, gap_open2=0, gap_ext2=0)))
    out.append(("affine_edge_shapes", PoaBatch.from_strings([["ACGTACGT", "ACGTTCGT", "ACGACGT"], ["A"], ["AC", "", "G"], [], ["NNNNACGTNN", "ACGT", "NNNN"]]), P(out_msa=True, **AFF)))
    out.append(("linear_edge_shapes", PoaBatch.from_strings([["ACGTACGT", "ACGTTCGT", "ACGACGT"], ["A"], ["AC", "", "G"], [], ["NNNNACGTNN", "ACGT", "NNNN"]]), P(out_msa=True, **LIN)))
    return out


def main():
    ref = RefAbpoa()
    store = {}
    names = []
    for name, batch, kw in cases():
        p = make_params(**kw)
        dumps = ref.poa_batch(p, batch, instrument=True)
        names.append(name)
        store[f"{name}/bso"] = batch.block_seq_off; store[f"{name}/sl"] = batch.seq_len; store[f"{name}/so"] = batch.seq_off
        store[f"{name}/ba"] = batch.bases; store[f"{name}/wt"] = batch.weight
        store[f"{name}/params"] = np.array([p.match, p.mismatch, p.gap_open1, p.gap_ext1, p.gap_open2, p.gap_ext2, p.align_mode, p.wb,
                                           p.out_cons, p.out_msa], dtype=np.int32)
        for i, d in enumerate(dumps):
            store[f"{name}/dump{i}"] = d.raw
        print(name, batch.n_blocks, "blocks", [d.n_node for d in dumps])
    store["names"] = np.array(names)
    store["ref_simd"] = np.array([ref.simd])
    out = os.path.join(ROOT, "tests", "golden", "abpoa_golden.npz")
    np.savez_compressed(out, **store)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
