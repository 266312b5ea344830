"""Test helpers: turn the product's C-ABI block view into the oracle's canonical dump layout
(oracle/poa_dump.h) so parity is a single np.array_equal."""
from __future__ import annotations

import numpy as np

from oracle.oracle import (Dump, PD_HEADER_LEN, PD_MAGIC, PD_N_NODE, PD_N_SEQ, PD_CONS_LEN, PD_MSA_LEN, PD_MSA_ROWS,
                           PD_N_IN_TOT, PD_N_OUT_TOT, PD_N_ALN_TOT, PD_PATH_TOT, PD_CIGAR_TOT, PD_INBAND_LO,
                           PD_INBAND_HI, POA_DUMP_MAGIC)


def view_to_dump(v) -> Dump:
    assert v.status == 0, f"block status {v.status}"
    hdr = np.zeros(PD_HEADER_LEN, dtype=np.int32)
    hdr[PD_MAGIC] = POA_DUMP_MAGIC
    hdr[PD_N_NODE] = v.n_node; hdr[PD_N_SEQ] = v.n_seq
    hdr[PD_CONS_LEN] = v.cons_len; hdr[PD_MSA_LEN] = v.msa_len; hdr[PD_MSA_ROWS] = v.msa_rows
    hdr[PD_N_IN_TOT] = v.in_id.shape[0]; hdr[PD_N_OUT_TOT] = v.out_id.shape[0]; hdr[PD_N_ALN_TOT] = v.aln_id.shape[0]
    hdr[PD_PATH_TOT] = v.path_node.shape[0]; hdr[PD_CIGAR_TOT] = v.cigar.shape[0]
    hdr[PD_INBAND_LO] = np.uint32(v.inband_cells & 0xFFFFFFFF).astype(np.int32)
    hdr[PD_INBAND_HI] = np.uint32(v.inband_cells >> 32).astype(np.int32)
    cig = np.zeros(2 * v.cigar.shape[0], dtype=np.int32)
    if v.cigar.shape[0]:
        cig[0::2] = (v.cigar & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.int32)
        cig[1::2] = (v.cigar >> np.uint64(32)).astype(np.uint32).view(np.int32)
    parts = [hdr, v.base, v.in_n, v.in_id, v.in_w, v.out_n, v.out_id, v.out_w, v.aln_n, v.aln_id,
             v.path_len, v.path_node, v.cons_node, v.msa.astype(np.int32), v.best_score, v.n_cigar, cig]
    return Dump(np.concatenate([np.asarray(p, dtype=np.int32) for p in parts]))


def first_diff(a: Dump, b: Dump) -> str:
    sa, sb = a.sections(), b.sections()
    for k in sa:
        if sa[k].shape != sb[k].shape:
            return f"section {k}: shape {sa[k].shape} vs {sb[k].shape}"
        if not np.array_equal(sa[k], sb[k]):
            i = int(np.nonzero(sa[k] != sb[k])[0][0])
            return f"section {k}[{i}]: {sa[k][i]} vs {sb[k][i]}"
    if not np.array_equal(a.raw[:PD_HEADER_LEN], b.raw[:PD_HEADER_LEN]):
        return f"header {a.raw[:PD_HEADER_LEN]} vs {b.raw[:PD_HEADER_LEN]}"
    return "equal"
