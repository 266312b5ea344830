/* TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
 *
 * Shim over the UNMODIFIED vendored abPOA (compiled from /root/reference/deps/abPOA by
 * oracle/Makefile).  It drives abPOA exactly the way smoothxg's smooth_abpoa does
 * (src/smooth.cpp:256-351): same parameter block, abpoa_reset(ab, abpt, 1024), one
 * constant per-base weight per sequence (the dedup multiplicity), abpoa_poa, then
 * abpoa_generate_rc_msa / abpoa_generate_consensus when requested -- and serialises what the
 * host side of smoothxg later reads from abpoa_t into the canonical dump of poa_dump.h.
 *
 * Two drive modes:
 *   instrument = 0 : calls abpoa_poa() itself (deps/abPOA/src/abpoa_align.c:304), i.e. the exact call smoothxg makes.
 *   instrument = 1 : replays abpoa_poa's loop (amb_strand = 0, so it is align + add) to also record
 *                    per-sequence best score, cigar and the in-band cell count from abm->dp_beg/dp_end.
 * ref_poa_batch_timed() is the CPU baseline: an OpenMP schedule(dynamic,1) loop over blocks like
 * src/smooth.cpp:1931, one abpoa_t/abpoa_para_t per block.
 */
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <omp.h>
#include "abpoa.h"
#include "poa_dump.h"

/* non-static symbols of the vendored library that abpoa.h does not declare */
extern int abpoa_poa(abpoa_t *ab, abpoa_para_t *abpt, uint8_t **seqs, int **weights, int *seq_lens, int exist_n_seq, int n_seq);
extern abpoa_seq_t *abpoa_realloc_seq(abpoa_seq_t *abs);

const char *ref_simd_name(void) {
#if __AVX512BW__
    return "avx512bw";
#elif defined(__AVX2__)
    return "avx2";
#elif defined(__SSE4_1__)
    return "sse4.1";
#else
    return "sse2";
#endif
}

void ref_free(void *p) { free(p); }

static abpoa_para_t *make_para(const pd_params_t *p) {
    abpoa_para_t *abpt = abpoa_init_para();
    abpt->align_mode = p->align_mode ? ABPOA_LOCAL_MODE : ABPOA_GLOBAL_MODE;
    abpt->wb = p->wb;
    abpt->wf = p->wf;
    abpt->amb_strand = 0;
    abpt->rev_cigar = 0;
    abpt->out_cons = p->out_cons ? 1 : 0;
    abpt->out_gfa = 1;
    abpt->out_msa = p->out_msa ? 1 : 0;
    abpt->match = p->match;
    abpt->mismatch = p->mismatch;
    abpt->gap_open1 = p->gap_open1;
    abpt->gap_open2 = p->gap_open2;
    abpt->gap_ext1 = p->gap_ext1;
    abpt->gap_ext2 = p->gap_ext2;
    abpt->disable_seeding = 1;
    abpt->k = 19; abpt->w = 10; abpt->min_w = 3313;
    abpoa_post_set_para(abpt);
    return abpt;
}

static int ilog2_u64(uint64_t v) { int r = 0; while (v >>= 1) ++r; return r; }

typedef struct {
    int64_t inband, full, edge_rows;
    int32_t *best_score, *n_cigar;
    pd_buf_t cigar;
} instr_t;

/* run one block; returns ab with the final graph (caller frees) */
static abpoa_t *run_block(abpoa_para_t *abpt, int n_seq, const int32_t *seq_len, const uint8_t *bases,
                          const int32_t *weight, instr_t *ins) {
    abpoa_t *ab = abpoa_init();
    int i, j;
    int *seq_lens = (int*)malloc(sizeof(int) * n_seq);
    uint8_t **bseqs = (uint8_t**)malloc(sizeof(uint8_t*) * n_seq);
    int **w = (int**)malloc(sizeof(int*) * n_seq);
    int64_t off = 0;
    for (i = 0; i < n_seq; ++i) {
        seq_lens[i] = seq_len[i];
        bseqs[i] = (uint8_t*)malloc(seq_len[i] > 0 ? seq_len[i] : 1);
        memcpy(bseqs[i], bases + off, seq_len[i]);
        off += seq_len[i];
        w[i] = (int*)malloc(sizeof(int) * (seq_len[i] > 0 ? seq_len[i] : 1));
        for (j = 0; j < seq_len[i]; ++j) w[i][j] = weight[i];
    }
    abpoa_reset(ab, abpt, 1024);
    abpoa_seq_t *abs = ab->abs; int exist_n_seq = abs->n_seq;
    abs->n_seq += n_seq; abpoa_realloc_seq(abs);
    for (i = 0; i < n_seq; ++i) { abs->name[exist_n_seq+i].l = 0; abs->name[exist_n_seq+i].m = 0; }

    if (!ins) {
        abpoa_poa(ab, abpt, bseqs, w, seq_lens, exist_n_seq, n_seq);
    } else {
        for (i = 0; i < n_seq; ++i) {
            abpoa_res_t res; res.graph_cigar = 0; res.n_cigar = 0; res.best_score = 0;
            int qlen = seq_lens[i];
            int gn = ab->abg->node_n;
            if (abpoa_align_sequence_to_graph(ab, abpt, bseqs[i], qlen, &res) >= 0) {
                int r;
                for (r = 0; r < gn - 1; ++r) {
                    int wdt = ab->abm->dp_end[r] - ab->abm->dp_beg[r] + 1;
                    ins->inband += wdt;
                    if (r > 0) ins->edge_rows += (int64_t)wdt * ab->abg->node[ab->abg->index_to_node_id[r]].in_edge_n;
                }
                ins->full += (int64_t)(gn - 1) * (qlen + 1);
                ins->best_score[i] = res.best_score;
                ins->n_cigar[i] = res.n_cigar;
                for (j = 0; j < res.n_cigar; ++j) {
                    pd_push(&ins->cigar, (int32_t)(uint32_t)(res.graph_cigar[j] & 0xffffffffULL));
                    pd_push(&ins->cigar, (int32_t)(uint32_t)(res.graph_cigar[j] >> 32));
                }
            }
            abpoa_add_graph_alignment(ab, abpt, bseqs[i], w[i], qlen, NULL, res, exist_n_seq + i, exist_n_seq + n_seq, 1);
            if (res.n_cigar) free(res.graph_cigar);
        }
    }
    /* src/smooth.cpp:342-351 */
    if (abpt->out_msa) abpoa_generate_rc_msa(ab, abpt);
    if (abpt->out_cons) abpoa_generate_consensus(ab, abpt);

    for (i = 0; i < n_seq; ++i) { free(bseqs[i]); free(w[i]); }
    free(bseqs); free(w); free(seq_lens);
    return ab;
}

static int32_t *dump_block(abpoa_t *ab, abpoa_para_t *abpt, int n_seq, instr_t *ins, int64_t *n_out) {
    abpoa_graph_t *abg = ab->abg;
    pd_buf_t b = {0, 0, 0};
    int i, j, k;
    for (i = 0; i < PD_HEADER_LEN; ++i) pd_push(&b, 0);
    int n_node = abg->node_n;
    int64_t n_in = 0, n_out_e = 0, n_aln = 0;
    for (i = 0; i < n_node; ++i) pd_push(&b, abg->node[i].base);
    for (i = 0; i < n_node; ++i) { pd_push(&b, abg->node[i].in_edge_n); n_in += abg->node[i].in_edge_n; }
    for (i = 0; i < n_node; ++i) for (j = 0; j < abg->node[i].in_edge_n; ++j) pd_push(&b, abg->node[i].in_id[j]);
    for (i = 0; i < n_node; ++i) for (j = 0; j < abg->node[i].in_edge_n; ++j) pd_push(&b, abg->node[i].in_edge_weight[j]);
    for (i = 0; i < n_node; ++i) { pd_push(&b, abg->node[i].out_edge_n); n_out_e += abg->node[i].out_edge_n; }
    for (i = 0; i < n_node; ++i) for (j = 0; j < abg->node[i].out_edge_n; ++j) pd_push(&b, abg->node[i].out_id[j]);
    for (i = 0; i < n_node; ++i) for (j = 0; j < abg->node[i].out_edge_n; ++j) pd_push(&b, abg->node[i].out_edge_weight[j]);
    for (i = 0; i < n_node; ++i) { pd_push(&b, abg->node[i].aligned_node_n); n_aln += abg->node[i].aligned_node_n; }
    for (i = 0; i < n_node; ++i) for (j = 0; j < abg->node[i].aligned_node_n; ++j) pd_push(&b, abg->node[i].aligned_node_id[j]);

    /* read paths: the same Kahn walk + bitset decode as build_odgi_abPOA, src/smooth.cpp:2457-2510 */
    int **paths = (int**)malloc(sizeof(int*) * n_seq);
    int *plen = (int*)calloc(n_seq, sizeof(int));
    for (i = 0; i < n_seq; ++i) paths[i] = (int*)malloc(sizeof(int) * (n_node > 0 ? n_node : 1));
    if (n_node > 2) {
        int *indeg = (int*)malloc(sizeof(int) * n_node);
        int *queue = (int*)malloc(sizeof(int) * n_node);
        int qh = 0, qt = 0;
        for (i = 0; i < n_node; ++i) indeg[i] = abg->node[i].in_edge_n;
        queue[qt++] = ABPOA_SRC_NODE_ID;
        while (qh < qt) {
            int cur = queue[qh++];
            if (cur == ABPOA_SINK_NODE_ID) break;
            if (cur != ABPOA_SRC_NODE_ID) {
                int base_id = 0;
                for (k = 0; k < abg->node[cur].read_ids_n; ++k) {
                    for (j = 0; j < abg->node[cur].out_edge_n; ++j) {
                        uint64_t num = abg->node[cur].read_ids[j][k];
                        while (num) {
                            uint64_t tmp = num & -num;
                            int rid = base_id + ilog2_u64(tmp);
                            paths[rid][plen[rid]++] = cur;
                            num ^= tmp;
                        }
                    }
                    base_id += 64;
                }
            }
            for (j = 0; j < abg->node[cur].out_edge_n; ++j) {
                int o = abg->node[cur].out_id[j];
                if (--indeg[o] == 0) queue[qt++] = o;
            }
        }
        free(indeg); free(queue);
    }
    int64_t path_tot = 0;
    for (i = 0; i < n_seq; ++i) { pd_push(&b, plen[i]); path_tot += plen[i]; }
    for (i = 0; i < n_seq; ++i) for (j = 0; j < plen[i]; ++j) pd_push(&b, paths[i][j]);
    for (i = 0; i < n_seq; ++i) free(paths[i]);
    free(paths); free(plen);

    int cons_len = -1, msa_len = -1, msa_rows = 0;
    abpoa_cons_t *abc = ab->abc;
    if (abpt->out_cons) {
        cons_len = (abg->is_called_cons && abc->n_cons > 0) ? abc->cons_len[0] : 0;
        for (i = 0; i < cons_len; ++i) pd_push(&b, abc->cons_node_ids[0][i]);
    }
    if (abpt->out_msa) {
        msa_len = abc->msa_len;
        msa_rows = n_seq + ((abpt->out_cons && abc->n_cons > 0) ? 1 : 0);
        if (msa_len <= 0) { msa_len = 0; msa_rows = 0; }
        for (i = 0; i < msa_rows; ++i) for (j = 0; j < msa_len; ++j) pd_push(&b, abc->msa_base[i][j]);
    }
    int64_t cig_tot = 0;
    for (i = 0; i < n_seq; ++i) pd_push(&b, ins ? ins->best_score[i] : 0);
    for (i = 0; i < n_seq; ++i) pd_push(&b, ins ? ins->n_cigar[i] : 0);
    if (ins) {
        cig_tot = ins->cigar.n / 2;
        for (int64_t q = 0; q < ins->cigar.n; ++q) pd_push(&b, ins->cigar.d[q]);
    }
    b.d[PD_MAGIC] = POA_DUMP_MAGIC;
    b.d[PD_N_NODE] = n_node; b.d[PD_N_SEQ] = n_seq;
    b.d[PD_CONS_LEN] = cons_len; b.d[PD_MSA_LEN] = msa_len; b.d[PD_MSA_ROWS] = msa_rows;
    b.d[PD_N_IN_TOT] = (int32_t)n_in; b.d[PD_N_OUT_TOT] = (int32_t)n_out_e; b.d[PD_N_ALN_TOT] = (int32_t)n_aln;
    b.d[PD_PATH_TOT] = (int32_t)path_tot; b.d[PD_CIGAR_TOT] = (int32_t)cig_tot;
    if (ins) {
        pd_set64(b.d, PD_INBAND_LO, ins->inband); pd_set64(b.d, PD_FULL_LO, ins->full);
        pd_set64(b.d, PD_EDGE_ROWS_LO, ins->edge_rows);
    }
    *n_out = b.n;
    return b.d;
}

int32_t *ref_poa_block(const pd_params_t *p, int n_seq, const int32_t *seq_len, const uint8_t *bases,
                       const int32_t *weight, int instrument, int64_t *n_out) {
    abpoa_para_t *abpt = make_para(p);
    instr_t ins; memset(&ins, 0, sizeof(ins));
    if (instrument) {
        ins.best_score = (int32_t*)calloc(n_seq > 0 ? n_seq : 1, sizeof(int32_t));
        ins.n_cigar = (int32_t*)calloc(n_seq > 0 ? n_seq : 1, sizeof(int32_t));
    }
    abpoa_t *ab = run_block(abpt, n_seq, seq_len, bases, weight, instrument ? &ins : NULL);
    int32_t *d = dump_block(ab, abpt, n_seq, instrument ? &ins : NULL, n_out);
    if (instrument) { free(ins.best_score); free(ins.n_cigar); free(ins.cigar.d); }
    abpoa_free(ab); abpoa_free_para(abpt);
    return d;
}

static uint64_t fnv1a(uint64_t h, const void *data, size_t n) {
    const uint8_t *p = (const uint8_t*)data; size_t i;
    for (i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ULL; }
    return h;
}

/* CPU baseline: time abpoa over a batch of blocks with n_threads OpenMP threads.
 * Returns wall seconds around the parallel loop only. per_block_hash (optional, n_blocks entries)
 * receives an FNV-1a hash over (n_node, bases, out ids, out weights) for cross-checking. */
double ref_poa_batch_timed(const pd_params_t *p, int n_blocks, const int64_t *block_seq_off,
                           const int32_t *seq_len, const int64_t *seq_off, const uint8_t *bases,
                           const int32_t *weight, int n_threads, uint64_t *per_block_hash) {
    if (n_threads <= 0) n_threads = omp_get_max_threads();
    struct timespec t0, t1;
    /* touch the global tables once so threads only ever rewrite identical values */
    { abpoa_para_t *w = make_para(p); abpoa_free_para(w); }
    clock_gettime(CLOCK_MONOTONIC, &t0);
    #pragma omp parallel for schedule(dynamic,1) num_threads(n_threads)
    for (int bi = 0; bi < n_blocks; ++bi) {
        int64_t s0 = block_seq_off[bi], s1 = block_seq_off[bi+1];
        int n_seq = (int)(s1 - s0);
        abpoa_para_t *abpt = make_para(p);
        abpoa_t *ab = run_block(abpt, n_seq, seq_len + s0, bases + seq_off[s0], weight + s0, NULL);
        if (per_block_hash) {
            uint64_t h = 1469598103934665603ULL; int i;
            abpoa_graph_t *abg = ab->abg;
            h = fnv1a(h, &abg->node_n, sizeof(int));
            for (i = 0; i < abg->node_n; ++i) {
                h = fnv1a(h, &abg->node[i].base, 1);
                h = fnv1a(h, abg->node[i].out_id, sizeof(int) * abg->node[i].out_edge_n);
                h = fnv1a(h, abg->node[i].out_edge_weight, sizeof(int) * abg->node[i].out_edge_n);
            }
            per_block_hash[bi] = h;
        }
        abpoa_free(ab); abpoa_free_para(abpt);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
