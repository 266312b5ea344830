#!/usr/bin/env python
"""Generate tests/golden/drb1_golden.npz: REAL POA blocks harvested from the reference's own test data.

How the inputs were harvested (reference src/smooth.cpp:527-536 is the reference's own dump hook):
  1. build the reference from a writable copy with -DPOA_DEBUG=ON (recipe: integration/build_reference.sh);
     integration/harvest.patch (26 lines, harvest tooling only, the algorithm is untouched) makes the POA_DEBUG FASTA
     header also carry what the FASTA otherwise loses -- the dedup weight, the per-duplicate strand flags and names --
     and re-enables the reference's commented-out dump of the block graph smooth_abpoa returns (src/smooth.cpp:622-624);
  2. run the ctest input (CMakeLists.txt:562-567) through the abPOA path, one target length:
        smoothxg -t 2 -g test/data/DRB1-3123...seqwish.gfa -j 5k -e 5k -l 1100 -r 12 -A -Z -B 0   (global, banded)
        smoothxg ... -A -B 0                                                                          (local)
     -> 17 files smoothxg_into_abpoa_pad311_<block>_in_<ms>ms.fa and 17 smoothxg_abpoa_block_<block>_final.gfa per mode;
  3. this script: python tests/golden/make_real_golden.py /tmp/harv_g /tmp/harv_l

For every block the npz stores the exact sequences (codes), dedup weights, padding, names / strand flags, the canonical
dump (oracle/poa_dump.h) of the UNMODIFIED vendored abPOA (oracle/_ref) run the way smooth_abpoa drives it, and the text
of the block graph the reference returned (checker of the graph-emission rows, tests/test_final_graph.py).
Real blocks are what the synthetic fixtures cannot be: N padding at path ends, dedup weights > 1, predecessor edges
hundreds of rows long, in-degree up to 6.
"""
import glob
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import RefAbpoa, make_params  # noqa: E402
from smoothxg_b200.synth import PoaBatch, encode  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "drb1_golden.npz")


def read_block(path):
    recs = []
    name = None
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith(">"):
            name = line[1:].rsplit(" ", 1)[0]
        elif line or name is not None:
            m = re.match(r"(.*);w=(\d+);revs=([01]+);names=(.*)$", name)
            recs.append(dict(first=m.group(1), w=int(m.group(2)), revs=m.group(3), names=m.group(4).split(","), seq=line))
            name = None
    return recs


def main(dirs):
    ref = RefAbpoa()
    out = {"names": []}
    for d, (mode, pk) in zip(dirs, (("global", dict()), ("local", dict(local=True)))):
        files = glob.glob(os.path.join(d, "smoothxg_into_abpoa_pad*_in_*ms.fa"))
        byid = {}
        for f in files:
            m = re.search(r"_pad(\d+)_(\d+)_in_", os.path.basename(f))
            byid[int(m.group(2))] = (f, int(m.group(1)))
        blocks, meta, finals = [], [], []
        for bid in sorted(byid):
            f, pad = byid[bid]
            recs = read_block(f)
            blocks.append(([encode(r["seq"]) for r in recs], [r["w"] for r in recs]))
            meta.append(f"{bid}\t{pad}\t" + "\t".join(f"{r['w']}:{r['revs']}:{','.join(r['names'])}" for r in recs))
            finals.append(open(os.path.join(d, f"smoothxg_abpoa_block_{bid}_final.gfa")).read())
        batch = PoaBatch.from_blocks(blocks)
        p = make_params(out_cons=True, out_msa=False, **pk)
        name = f"drb1_{mode}"
        out["names"].append(name)
        out[f"{name}/bso"] = batch.block_seq_off; out[f"{name}/sl"] = batch.seq_len; out[f"{name}/so"] = batch.seq_off
        out[f"{name}/ba"] = batch.bases; out[f"{name}/wt"] = batch.weight
        out[f"{name}/params"] = np.array([p.match, p.mismatch, p.gap_open1, p.gap_ext1, p.gap_open2, p.gap_ext2, p.align_mode, p.wb, p.out_cons, p.out_msa], dtype=np.int32)
        out[f"{name}/meta"] = np.array(meta)
        out[f"{name}/final_gfa"] = np.array(finals)
        cells = 0
        for b in range(batch.n_blocks):
            dmp = ref.poa_block(p, *batch.block(b), instrument=True)
            out[f"{name}/dump{b}"] = dmp.raw
            cells += dmp.inband_cells
        print(f"{name}: {batch.n_blocks} blocks, {batch.n_seqs} sequences, {batch.bases.shape[0]} bases, {cells} in-band cells ({ref.simd})")
    out["names"] = np.array(out["names"])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main(sys.argv[1:3])
