/* TEST INFRASTRUCTURE ONLY.
 *
 * Canonical int32 "dump" of one POA block result, shared by the two checkers
 * (ref_shim.c = unmodified vendored abPOA, poa_oracle.c = scalar restatement) and
 * mirrored in Python by oracle/oracle.py (which builds the same dump from the
 * product's C-ABI output so tests can compare with np.array_equal).
 *
 * Everything the host side of smoothxg reads from abpoa_t after abpoa_poa
 * (src/smooth.cpp:342-516, build_odgi_abPOA src/smooth.cpp:2442-2574) is in here:
 * node bases, in/out edge lists in their final (weight-sorted) order with weights,
 * per-read node paths (decoded from the read_ids bitsets exactly like
 * src/smooth.cpp:2488-2501), consensus node ids, and the RC-MSA.
 */
#ifndef POA_DUMP_H
#define POA_DUMP_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define POA_DUMP_MAGIC 0x504f4131 /* "POA1" */

enum {
    PD_MAGIC = 0,
    PD_N_NODE,
    PD_N_SEQ,
    PD_CONS_LEN,   /* -1 when no consensus requested */
    PD_MSA_LEN,    /* -1 when no msa requested */
    PD_MSA_ROWS,
    PD_N_IN_TOT,
    PD_N_OUT_TOT,
    PD_N_ALN_TOT,
    PD_PATH_TOT,
    PD_CIGAR_TOT,  /* number of 64-bit cigar words over all sequences (0 if not recorded) */
    PD_INBAND_LO, PD_INBAND_HI,  /* in-band DP cells, sum over aligned sequences */
    PD_FULL_LO, PD_FULL_HI,      /* full-matrix equivalent cells: sum rows*(qlen+1) */
    PD_EDGE_ROWS_LO, PD_EDGE_ROWS_HI, /* sum over evaluated rows of predecessor count * band width (for p-bar) */
    PD_HEADER_LEN = 24
};

/* Sections after the header, in order:
 *   base[n_node]
 *   in_n[n_node]  in_id[n_in_tot]  in_w[n_in_tot]
 *   out_n[n_node] out_id[n_out_tot] out_w[n_out_tot]
 *   aln_n[n_node] aln_id[n_aln_tot]
 *   path_len[n_seq] path_node[path_tot]
 *   cons_node[max(cons_len,0)]
 *   msa[msa_rows*max(msa_len,0)]
 *   best_score[n_seq] n_cigar[n_seq]
 *   cigar[2*cigar_tot]  (lo32, hi32 of each abpoa_cigar_t word)
 */

typedef struct {
    int32_t *d; int64_t n, m;
} pd_buf_t;

static inline void pd_push(pd_buf_t *b, int32_t v) {
    if (b->n == b->m) {
        b->m = b->m ? b->m * 2 : 4096;
        b->d = (int32_t*)realloc(b->d, (size_t)b->m * sizeof(int32_t));
    }
    b->d[b->n++] = v;
}

static inline void pd_set64(int32_t *hdr, int lo_idx, int64_t v) {
    hdr[lo_idx] = (int32_t)(uint32_t)(v & 0xffffffffLL);
    hdr[lo_idx + 1] = (int32_t)(uint32_t)((uint64_t)v >> 32);
}

/* Parameter block handed to both checkers; mirrors include/poa_b200.h poa_b200_params_t
 * (same field order and meaning) so tests pass one struct to all three. */
typedef struct {
    int32_t match, mismatch, gap_open1, gap_ext1, gap_open2, gap_ext2; /* positive penalties, src/smooth.cpp:282-287 */
    int32_t align_mode;  /* 0 global, 1 local (src/smooth.cpp:259-263) */
    int32_t wb;          /* 311 banded / -1 unbanded (src/smooth.cpp:266-270) */
    float   wf;          /* 0.03 (src/smooth.cpp:271) */
    int32_t out_cons;    /* src/smooth.cpp:275 */
    int32_t out_msa;     /* src/smooth.cpp:278 */
} pd_params_t;

#endif
