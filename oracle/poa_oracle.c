/* TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path
 * (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this).
 *
 * Scalar C restatement of the abPOA v1.5.4 path that smoothxg's smooth_abpoa drives
 * (reference = /root/reference, citations relative to it):
 *   abpoa_poa                         deps/abPOA/src/abpoa_align.c:304-344
 *   simd_abpoa_align_sequence_to_subgraph (int16/int32 choice, inf_min)
 *                                     deps/abPOA/src/abpoa_align_simd.c:1250-1332
 *   convex-gap row recurrence         deps/abPOA/src/abpoa_align_simd.c:935-1074 (first row :617-688)
 *   adaptive band                     deps/abPOA/src/abpoa_align.h:34-35, abpoa_align_simd.c:1107-1130
 *   best cell                         deps/abPOA/src/abpoa_align_simd.c:1092-1105 (global), :1208-1210 (local)
 *   backtrack                         deps/abPOA/src/abpoa_align_simd.c:309-458
 *   graph fusion                      deps/abPOA/src/abpoa_graph.c:688-773, :480-556, :573-592, :450-463
 *   topological sort                  deps/abPOA/src/abpoa_graph.c:322-357 (:221-266, :192-219, :268-309)
 *   heaviest-bundle consensus         deps/abPOA/src/abpoa_output.c:468-536, :375-391
 *   RC-MSA                            deps/abPOA/src/abpoa_output.c:149-192; rank abpoa_graph.c:359-419
 *
 * No SIMD: the reference's striped vectors are restated cell by cell.  A cell
 * outside a predecessor row's [dp_beg,dp_end] reads as the finite inf_min; all score arithmetic wraps
 * in the chosen score width exactly like _mm*_add/sub_epi16/32.  The one place the reference's lane
 * count pn leaks into observable state is the band start (abpoa_align_simd.c:957-959: beg is raised
 * to the predecessors' minimum only if it lies in an earlier *vector*); `pn16`/`pn32` reproduce
 * that (default 32/16 = the AVX-512BW build).  It only moves junk (-inf) cells, so outputs do not
 * depend on it; the in-band cell count does.
 *
 * Parity pinning: upstream has no known-answer tests for abPOA.  This restatement is
 * pinned by bit-comparing its canonical dump (poa_dump.h) with the unmodified vendored abPOA run
 * through oracle/ref_shim.c on every fixture (tests/test_oracle_vs_ref.py, tests/golden/).
 *
 * Scope: all three gap modes of abpoa_set_gap_mode (abpoa_align.c:87-91) -- convex (smoothxg's default
 * 1,4,6,2,26,1 and all five adaptive presets, src/smooth.cpp:2028-2062), affine (four-value -p,
 * src/main.cpp:353-359) and linear -- global (banded or not) and local alignment.
 */
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include "poa_dump.h"

#define SRC_ID 0
#define SINK_ID 1
#define OP_M  0x1
#define OP_E1 0x2
#define OP_E2 0x4
#define OP_E  0x6
#define OP_F1 0x8
#define OP_F2 0x10
#define OP_F  0x18
#define OP_ALL 0x1f
#define CMATCH 0
#define CINS 1
#define CDEL 2

static int g_pn16 = 32, g_pn32 = 16;
void oracle_set_lane_counts(int pn16, int pn32) { g_pn16 = pn16; g_pn32 = pn32; }
void oracle_free(void *p) { free(p); }

typedef struct { int *id, *w; int n, m; } elist_t;
typedef struct { uint8_t base; elist_t in, out; int aln[8]; int aln_n; } node_t;
typedef struct {
    node_t *node; int n, m;
    int *idx2id, *id2idx, *mpl, *mpr, *remain;
    int arr_m;
} graph_t;

static void elist_push(elist_t *e, int id, int w) {
    if (e->n == e->m) {
        e->m = e->m ? e->m * 2 : 2;
        e->id = (int*)realloc(e->id, sizeof(int) * e->m);
        e->w = (int*)realloc(e->w, sizeof(int) * e->m);
    }
    e->id[e->n] = id; e->w[e->n] = w; e->n++;
}

static int add_node(graph_t *g, uint8_t base) { /* abpoa_graph.c:471-478 */
    if (g->n == g->m) {
        int om = g->m;
        g->m = g->m ? g->m * 2 : 1024;
        g->node = (node_t*)realloc(g->node, sizeof(node_t) * g->m);
        memset(g->node + om, 0, sizeof(node_t) * (g->m - om));
    }
    g->node[g->n].base = base;
    return g->n++;
}

/* abpoa_graph.c:480-556 (read ids are recorded as explicit per-read paths by the caller) */
static void add_edge(graph_t *g, int from, int to, int check_edge, int w) {
    int exist = 0, i;
    if (check_edge) {
        elist_t *in = &g->node[to].in;
        for (i = 0; i < in->n; ++i) if (in->id[i] == from) { in->w[i] += w; break; }
        elist_t *out = &g->node[from].out;
        for (i = 0; i < out->n; ++i) if (out->id[i] == to) { out->w[i] += w; exist = 1; break; }
    }
    if (!exist) {
        elist_push(&g->node[to].in, from, w);
        elist_push(&g->node[from].out, to, w);
    }
}

static void add_aligned1(node_t *nd, int id) { nd->aln[nd->aln_n++] = id; }
static void add_aligned(graph_t *g, int node_id, int new_id) { /* abpoa_graph.c:455-463 */
    int i; node_t *node = g->node;
    for (i = 0; i < node[node_id].aln_n; ++i) {
        add_aligned1(&node[node[node_id].aln[i]], new_id);
        add_aligned1(&node[new_id], node[node_id].aln[i]);
    }
    add_aligned1(&node[node_id], new_id);
    add_aligned1(&node[new_id], node_id);
}
static int get_aligned_id(graph_t *g, int node_id, uint8_t base) { /* abpoa_graph.c:439-448 */
    int i;
    for (i = 0; i < g->node[node_id].aln_n; ++i) {
        int a = g->node[node_id].aln[i];
        if (g->node[a].base == base) return a;
    }
    return -1;
}

/* abpoa_graph.c:221-266 */
static void bfs_set_node_index(graph_t *g) {
    int n = g->n, i, j, index = 0;
    int *indeg = (int*)malloc(sizeof(int) * n);
    int *q = (int*)malloc(sizeof(int) * (n + 8)); int qh = 0, qt = 0;
    for (i = 0; i < n; ++i) indeg[i] = g->node[i].in.n;
    q[qt++] = SRC_ID;
    while (qh < qt) {
        int cur = q[qh++];
        g->idx2id[index] = cur; g->id2idx[cur] = index++;
        if (cur == SINK_ID) break;
        for (i = 0; i < g->node[cur].out.n; ++i) {
            int o = g->node[cur].out.id[i];
            if (--indeg[o] == 0) {
                int ok = 1;
                for (j = 0; j < g->node[o].aln_n; ++j) if (indeg[g->node[o].aln[j]] != 0) { ok = 0; break; }
                if (!ok) continue;
                q[qt++] = o;
                for (j = 0; j < g->node[o].aln_n; ++j) q[qt++] = g->node[o].aln[j];
            }
        }
    }
    free(indeg); free(q);
}

/* abpoa_graph.c:192-219: exchange sort, strict <, NOT stable */
static void sort_in_out_ids(graph_t *g) {
    int i, j, k, t;
    for (i = 0; i < g->n; ++i) {
        elist_t *e = &g->node[i].in;
        for (j = 0; j < e->n - 1; ++j) for (k = j + 1; k < e->n; ++k) if (e->w[j] < e->w[k]) {
            t = e->id[j]; e->id[j] = e->id[k]; e->id[k] = t;
            t = e->w[j]; e->w[j] = e->w[k]; e->w[k] = t;
        }
        e = &g->node[i].out;
        for (j = 0; j < e->n - 1; ++j) for (k = j + 1; k < e->n; ++k) if (e->w[j] < e->w[k]) {
            t = e->id[j]; e->id[j] = e->id[k]; e->id[k] = t;
            t = e->w[j]; e->w[j] = e->w[k]; e->w[k] = t;
        }
    }
}

/* abpoa_graph.c:268-309 */
static void bfs_set_node_remain(graph_t *g) {
    int n = g->n, i;
    int *outdeg = (int*)malloc(sizeof(int) * n);
    int *q = (int*)malloc(sizeof(int) * (n + 8)); int qh = 0, qt = 0;
    for (i = 0; i < n; ++i) { outdeg[i] = g->node[i].out.n; g->remain[i] = 0; }
    q[qt++] = SINK_ID; g->remain[SINK_ID] = -1;
    while (qh < qt) {
        int cur = q[qh++];
        if (cur != SINK_ID) {
            int max_w = -1, max_id = -1;
            for (i = 0; i < g->node[cur].out.n; ++i)
                if (g->node[cur].out.w[i] > max_w) { max_w = g->node[cur].out.w[i]; max_id = g->node[cur].out.id[i]; }
            g->remain[cur] = g->remain[max_id] + 1;
        }
        if (cur == SRC_ID) break;
        for (i = 0; i < g->node[cur].in.n; ++i) {
            int p = g->node[cur].in.id[i];
            if (--outdeg[p] == 0) q[qt++] = p;
        }
    }
    free(outdeg); free(q);
}

/* abpoa_graph.c:322-357 */
static void topological_sort(graph_t *g, int banded) {
    int n = g->n, i;
    if (n > g->arr_m) {
        g->arr_m = n * 2;
        g->idx2id = (int*)realloc(g->idx2id, sizeof(int) * g->arr_m);
        g->id2idx = (int*)realloc(g->id2idx, sizeof(int) * g->arr_m);
        g->mpl = (int*)realloc(g->mpl, sizeof(int) * g->arr_m);
        g->mpr = (int*)realloc(g->mpr, sizeof(int) * g->arr_m);
        g->remain = (int*)realloc(g->remain, sizeof(int) * g->arr_m);
    }
    bfs_set_node_index(g);
    sort_in_out_ids(g);
    if (banded) {
        for (i = 0; i < n; ++i) { g->mpr[i] = 0; g->mpl[i] = n; }
        bfs_set_node_remain(g);
    }
}

/* ---------------------------------------------------------------- alignment */
typedef struct {
    int32_t best_score; int n_cigar, m_cigar; uint64_t *cigar;
    int64_t inband, full, edge_rows;
} aln_t;

static void push_cigar(aln_t *a, int op, int len, int32_t node_id, int32_t query_id) { /* abpoa_align.h:54-73 */
    uint64_t l = (uint64_t)len;
    if (a->n_cigar == 0 || op != CINS || op != (int)(a->cigar[a->n_cigar - 1] & 0xf)) {
        if (a->n_cigar == a->m_cigar) {
            a->m_cigar = a->m_cigar ? a->m_cigar << 1 : 4;
            a->cigar = (uint64_t*)realloc(a->cigar, sizeof(uint64_t) * a->m_cigar);
        }
        uint64_t n_id = (uint64_t)(int64_t)node_id, q_id = (uint64_t)(int64_t)query_id;
        if (op == CMATCH) a->cigar[a->n_cigar++] = n_id << 34 | q_id << 4 | op;
        else if (op == CINS) a->cigar[a->n_cigar++] = q_id << 34 | l << 4 | op;
        else a->cigar[a->n_cigar++] = n_id << 34 | l << 4 | op;
    } else a->cigar[a->n_cigar - 1] += l << 4;
}

typedef struct {
    int32_t *mem; size_t mem_m;
    int64_t *off; int *beg, *end, *beg_sn; int rows_m;
} dpm_t;

#define WRAP(x) (bits16 ? (int32_t)(int16_t)(x) : (int32_t)(uint32_t)(int64_t)(x))
#define MAX2(a, b) ((a) > (b) ? (a) : (b))
#define MIN2(a, b) ((a) < (b) ? (a) : (b))

/* one sequence against the current graph; deps/abPOA/src/abpoa_align_simd.c:1250-1332 + the cg core :1201-1231 */
static void align_sequence(graph_t *g, const pd_params_t *P, const int mat[25], const uint8_t *query, int qlen,
                           dpm_t *dp, aln_t *res) {
    const int gn = g->n; /* end_index - beg_index + 1 */
    const int local = P->align_mode == 1;
    const int wb = local ? -1 : P->wb;  /* abpoa_align.c:158 */
    const int32_t e1 = P->gap_ext1, e2 = P->gap_ext2, o1 = P->gap_open1, o2 = P->gap_open2;
    const int32_t oe1 = o1 + e1, oe2 = o2 + e2;
    const int32_t match = P->match < 0 ? -P->match : P->match;
    const int32_t min_mis = P->mismatch < 0 ? -P->mismatch : P->mismatch; /* abpoa_align.c:14-24 */
    /* :1286-1302 */
    int len = qlen > gn ? qlen : gn;
    int64_t max_score = MAX2((int64_t)qlen * match, (int64_t)len * e1 + o1);
    int bits16 = max_score <= INT16_MAX - min_mis - oe1 - oe2;
    int64_t base_min = bits16 ? INT16_MIN : INT32_MIN;
    int32_t inf_min = (int32_t)(MAX2(MAX2(base_min + min_mis, base_min + oe1), base_min + oe2) + 512 * MAX2(e1, e2));
    int pn = bits16 ? g_pn16 : g_pn32;
    int w = wb < 0 ? qlen : wb + (int)(P->wf * qlen); /* :474 */
    int rows = gn - 1; /* sink row is never filled */
    int i, j, k;

    if (rows > dp->rows_m) {
        dp->rows_m = rows * 2;
        dp->off = (int64_t*)realloc(dp->off, sizeof(int64_t) * dp->rows_m);
        dp->beg = (int*)realloc(dp->beg, sizeof(int) * dp->rows_m);
        dp->end = (int*)realloc(dp->end, sizeof(int) * dp->rows_m);
        dp->beg_sn = (int*)realloc(dp->beg_sn, sizeof(int) * dp->rows_m);
    }
    size_t used = 0;
#define ROW_ALLOC(r, b, e) do { \
        size_t need = used + (size_t)5 * ((e) - (b) + 1); \
        if (need > dp->mem_m) { dp->mem_m = need * 2 + (1 << 20); dp->mem = (int32_t*)realloc(dp->mem, sizeof(int32_t) * dp->mem_m); } \
        dp->off[r] = (int64_t)used; dp->beg[r] = (b); dp->end[r] = (e); dp->beg_sn[r] = (b) / pn; used = need; } while (0)
#define PL(r, p) (dp->mem + dp->off[r] + (int64_t)(p) * (dp->end[r] - dp->beg[r] + 1) - dp->beg[r])

    int64_t inband = 0, edge_rows = 0;
    /* ---- first row, :617-688 */
    {
        int end0;
        if (wb >= 0) {
            g->mpl[SRC_ID] = g->mpr[SRC_ID] = 0;
            for (i = 0; i < g->node[SRC_ID].out.n; ++i) { int o = g->node[SRC_ID].out.id[i]; g->mpl[o] = g->mpr[o] = 1; }
            int r = qlen - (g->remain[SRC_ID] - g->remain[SINK_ID] - 1);
            end0 = MIN2(qlen, MAX2(g->mpr[SRC_ID], r) + w);
        } else end0 = qlen;
        ROW_ALLOC(0, 0, end0);
        int32_t *H = PL(0, 0), *E1 = PL(0, 1), *E2 = PL(0, 2), *F1 = PL(0, 3), *F2 = PL(0, 4);
        if (local) {
            for (j = 0; j <= end0; ++j) H[j] = E1[j] = E2[j] = F1[j] = F2[j] = 0;
        } else {
            for (j = 0; j <= end0; ++j) { H[j] = E1[j] = E2[j] = inf_min; }
            H[0] = 0; E1[0] = WRAP(-oe1); E2[0] = WRAP(-oe2); F1[0] = F2[0] = inf_min;
            for (j = 1; j <= end0; ++j) {
                F1[j] = WRAP(-o1 - e1 * j); F2[j] = WRAP(-o2 - e2 * j);
                H[j] = MAX2(F1[j], F2[j]);
            }
        }
        inband += end0 + 1;
    }
    int32_t best_score = inf_min; int best_i = 0, best_j = 0;
    /* ---- rows in index order, :1205-1221 */
    for (i = 1; i < rows; ++i) {
        int v = g->idx2id[i];
        const elist_t *in = &g->node[v].in;
        int beg, end;
        if (wb < 0) { beg = 0; end = qlen; }
        else { /* abpoa_align.h:34-35, abpoa_align_simd.c:946-960 */
            int r = qlen - (g->remain[v] - g->remain[SINK_ID] - 1);
            beg = MAX2(0, MIN2(g->mpl[v], r) - w);
            end = MIN2(qlen, MAX2(g->mpr[v], r) + w);
            int beg_sn = beg / pn, min_pre_beg = INT_MAX, min_pre_beg_sn = INT_MAX;
            for (k = 0; k < in->n; ++k) {
                int pi = g->id2idx[in->id[k]];
                if (min_pre_beg > dp->beg[pi]) { min_pre_beg = dp->beg[pi]; min_pre_beg_sn = dp->beg_sn[pi]; }
            }
            if (beg_sn < min_pre_beg_sn) beg = min_pre_beg;
        }
        if (end < beg) end = beg; /* defensive; does not occur with w >= 0 */
        ROW_ALLOC(i, beg, end);
        int32_t *H = PL(i, 0), *E1 = PL(i, 1), *E2 = PL(i, 2), *F1 = PL(i, 3), *F2 = PL(i, 4);
        const int *mrow = mat + 5 * g->node[v].base;
        inband += end - beg + 1; edge_rows += (int64_t)in->n * (end - beg + 1);
        /* M / E from predecessors, :966-1029 */
        for (j = beg; j <= end; ++j) { H[j] = inf_min; E1[j] = inf_min; E2[j] = inf_min; }
        for (k = 0; k < in->n; ++k) {
            int pi = g->id2idx[in->id[k]];
            const int32_t *pH = PL(pi, 0), *pE1 = PL(pi, 1), *pE2 = PL(pi, 2);
            int pb = dp->beg[pi], pe = dp->end[pi];
            int lo = MAX2(beg, pb + 1), hi = MIN2(end, pe + 1);
            for (j = lo; j <= hi; ++j) if (pH[j - 1] > H[j]) H[j] = pH[j - 1];
            if (local && beg == 0 && 0 > H[0]) H[0] = 0; /* :974 `first` = 0 in local mode */
            lo = MAX2(beg, pb); hi = MIN2(end, pe);
            for (j = lo; j <= hi; ++j) {
                if (pE1[j] > E1[j]) E1[j] = pE1[j];
                if (pE2[j] > E2[j]) E2[j] = pE2[j];
            }
        }
        /* H = M + profile (:1032-1034, profile column 0 is 0 :536), then F scan and H/E update (:1038-1073) */
        int32_t prevHh = inf_min, f1 = inf_min, f2 = inf_min;
        for (j = beg; j <= end; ++j) {
            int32_t s = j == 0 ? 0 : mrow[query[j - 1]];
            int32_t hm = WRAP(H[j] + s);
            int32_t hh = MAX2(MAX2(hm, E1[j]), E2[j]);
            f1 = MAX2(WRAP(prevHh - oe1), WRAP(f1 - e1));
            f2 = MAX2(WRAP(prevHh - oe2), WRAP(f2 - e2));
            F1[j] = f1; F2[j] = f2;
            int32_t h = MAX2(hh, MAX2(f1, f2));
            if (local) h = MAX2(h, 0);
            H[j] = h;
            int32_t ne1 = MAX2(WRAP(E1[j] - e1), WRAP(h - oe1));
            int32_t ne2 = MAX2(WRAP(E2[j] - e2), WRAP(h - oe2));
            if (local) { ne1 = MAX2(ne1, 0); ne2 = MAX2(ne2, 0); }
            E1[j] = ne1; E2[j] = ne2;
            prevHh = hh;
        }
        /* max in row, :1107-1119 */
        if (local || wb >= 0) {
            int32_t mx = inf_min; int left = -1, right = -1;
            for (j = beg; j <= end; ++j) {
                if (H[j] > mx) { mx = H[j]; left = right = j; }
                else if (H[j] == mx) right = j;
            }
            if (local && mx > best_score) { best_score = mx; best_i = i; best_j = left; } /* :1208-1210 */
            if (wb >= 0) { /* :1121-1130 */
                const elist_t *out = &g->node[v].out;
                for (k = 0; k < out->n; ++k) {
                    int o = out->id[k];
                    if (right + 1 > g->mpr[o]) g->mpr[o] = right + 1;
                    if (left + 1 < g->mpl[o]) g->mpl[o] = left + 1;
                }
            }
        }
    }
    /* global best, :1092-1105 */
    if (!local) {
        const elist_t *in = &g->node[SINK_ID].in;
        for (k = 0; k < in->n; ++k) {
            int pi = g->id2idx[in->id[k]];
            int e = qlen > dp->end[pi] ? dp->end[pi] : qlen;
            int32_t sc = PL(pi, 0)[e];
            if (sc > best_score) { best_score = sc; best_i = pi; best_j = e; }
        }
    }
    res->best_score = best_score;
    res->inband = inband; res->edge_rows = edge_rows; res->full = (int64_t)rows * (qlen + 1);

    /* ---- backtrack, :309-458 (put_gap_on_right = put_gap_at_end = 0, inc_path_score = 0) */
    {
        int cur_op = OP_ALL, hit, id, s;
        i = best_i; j = best_j; id = g->idx2id[i];
        if (best_j < qlen) push_cigar(res, CINS, qlen - best_j, -1, qlen - 1);
        while (i > 0 && j > 0) {
            const int32_t *H = PL(i, 0), *E1 = PL(i, 1), *E2 = PL(i, 2), *F1 = PL(i, 3), *F2 = PL(i, 4);
            if (local && H[j] == 0) break;
            const elist_t *in = &g->node[id].in;
            s = mat[5 * g->node[id].base + query[j - 1]]; hit = 0;
            if (cur_op & OP_M) {
                for (k = 0; k < in->n; ++k) {
                    int pi = g->id2idx[in->id[k]];
                    if (j - 1 < dp->beg[pi] || j - 1 > dp->end[pi]) continue;
                    if (WRAP(PL(pi, 0)[j - 1] + s) == H[j]) {
                        push_cigar(res, CMATCH, 1, id, j - 1);
                        i = pi; --j; id = g->idx2id[i]; hit = 1; cur_op = OP_ALL;
                        break;
                    }
                }
            }
            if (hit == 0 && (cur_op & OP_E)) {
                for (k = 0; k < in->n; ++k) {
                    int pi = g->id2idx[in->id[k]];
                    if (j < dp->beg[pi] || j > dp->end[pi]) continue;
                    const int32_t *pH = PL(pi, 0);
                    if (cur_op & OP_E1) {
                        const int32_t *pE1 = PL(pi, 1);
                        int cond = (cur_op & OP_M) ? (H[j] == pE1[j]) : (E1[j] == WRAP(pE1[j] - e1));
                        if (cond) {
                            if (WRAP(pH[j] - oe1) == pE1[j]) cur_op = OP_M | OP_F; else cur_op = OP_E1;
                            hit = 1; push_cigar(res, CDEL, 1, id, j - 1);
                            i = pi; id = g->idx2id[i];
                            break;
                        }
                    }
                    if (cur_op & OP_E2) {
                        const int32_t *pE2 = PL(pi, 2);
                        int cond = (cur_op & OP_M) ? (H[j] == pE2[j]) : (E2[j] == WRAP(pE2[j] - e2));
                        if (cond) {
                            if (WRAP(pH[j] - oe2) == pE2[j]) cur_op = OP_M | OP_F; else cur_op = OP_E2;
                            hit = 1; push_cigar(res, CDEL, 1, id, j - 1);
                            i = pi; id = g->idx2id[i];
                            break;
                        }
                    }
                }
            }
            if (hit == 0 && (cur_op & OP_F)) {
                /* H[j-1] may lie left of this row's band: the reference then reads whatever the striped row
                 * holds there; our rows are band-only, so read inf_min (it can only matter on junk cells) */
                int32_t hl = (j - 1 >= dp->beg[i]) ? H[j - 1] : inf_min;
                int32_t f1l = (j - 1 >= dp->beg[i]) ? F1[j - 1] : inf_min;
                int32_t f2l = (j - 1 >= dp->beg[i]) ? F2[j - 1] : inf_min;
                if (cur_op & OP_F1) {
                    if (!(cur_op & OP_M) || H[j] == F1[j]) {
                        if (WRAP(hl - oe1) == F1[j]) { cur_op = OP_M | OP_E; hit = 1; }
                        else if (WRAP(f1l - e1) == F1[j]) { cur_op = OP_F1; hit = 1; }
                    }
                }
                if (hit == 0 && (cur_op & OP_F2)) {
                    if (!(cur_op & OP_M) || H[j] == F2[j]) {
                        if (WRAP(hl - oe2) == F2[j]) { cur_op = OP_M | OP_E; hit = 1; }
                        else if (WRAP(f2l - e2) == F2[j]) { cur_op = OP_F2; hit = 1; }
                    }
                }
                if (hit == 1) { push_cigar(res, CINS, 1, id, j - 1); --j; }
            }
            if (hit == 0) { fprintf(stderr, "[poa_oracle] backtrack dead end at (%d,%d) cur_op=%d\n", i, j, cur_op); abort(); }
        }
        if (j > 0) push_cigar(res, CINS, j, -1, j - 1);
        for (k = 0; k < res->n_cigar >> 1; ++k) { /* reverse, abpoa_align.h:88-96 */
            uint64_t t = res->cigar[k]; res->cigar[k] = res->cigar[res->n_cigar - 1 - k]; res->cigar[res->n_cigar - 1 - k] = t;
        }
    }
#undef ROW_ALLOC
#undef PL
}

/* ---- affine (gap_open1 > 0, gap_open2 == 0) and linear (gap_open1 == 0) gap modes (abpoa_align.c:87-91).
 * smoothxg reaches the affine kernel when -p has four values (src/main.cpp:353-359).
 *   affine row  simd_abpoa_ag_dp  abpoa_align_simd.c:817-933, first row :645-658, backtrack :196-307
 *   linear row  simd_abpoa_lg_dp  abpoa_align_simd.c:727-815, first row :631-643, backtrack :116-194
 * Differences from the convex kernel that are restated here on purpose:
 *   - affine: F1 is fed by M + profile alone, not by max(M + profile, E1) (:908); the stored E1 is reset to
 *     inf_min (0 in local mode) where F1 strictly won the cell (:926,:930); local mode does not clamp E1;
 *   - affine: the first cell of a row's first vector gets F1 = (M + profile) - oe1 of the SAME column (:898,:908);
 *     it never wins a cell or feeds one, but it is stored, so it is restated (vector start = beg rounded down to pn);
 *   - linear: one plane; a cell is max over predecessors of (H_p[j-1] + s, H_p[j] - e1), then the row-wise
 *     running max with H[j-1] - e1; local mode clamps at 0 only after that pass (:813);
 *   - traceback order: affine M, E1, F1; linear M, deletion, insertion (put_gap_on_right = put_gap_at_end = 0). */
static void align_sequence_al(graph_t *g, const pd_params_t *P, const int mat[25], const uint8_t *query, int qlen,
                              dpm_t *dp, aln_t *res, int linear) {
    const int gn = g->n;
    const int local = P->align_mode == 1;
    const int wb = local ? -1 : P->wb;
    const int32_t e1 = P->gap_ext1, e2 = P->gap_ext2, o1 = P->gap_open1, o2 = P->gap_open2;
    const int32_t oe1 = o1 + e1, oe2 = o2 + e2;
    const int32_t match = P->match < 0 ? -P->match : P->match;
    const int32_t min_mis = P->mismatch < 0 ? -P->mismatch : P->mismatch;
    int len = qlen > gn ? qlen : gn;
    int64_t max_score = MAX2((int64_t)qlen * match, (int64_t)len * e1 + o1);
    int bits16 = max_score <= INT16_MAX - min_mis - oe1 - oe2;
    int64_t base_min = bits16 ? INT16_MIN : INT32_MIN;
    int32_t inf_min = (int32_t)(MAX2(MAX2(base_min + min_mis, base_min + oe1), base_min + oe2) + 512 * MAX2(e1, e2));
    int pn = bits16 ? g_pn16 : g_pn32;
    int w = wb < 0 ? qlen : wb + (int)(P->wf * qlen);
    int rows = gn - 1;
    const int NPL = linear ? 1 : 3; /* H | H, E1, F1 */
    int i, j, k;
    if (rows > dp->rows_m) {
        dp->rows_m = rows * 2;
        dp->off = (int64_t*)realloc(dp->off, sizeof(int64_t) * dp->rows_m);
        dp->beg = (int*)realloc(dp->beg, sizeof(int) * dp->rows_m);
        dp->end = (int*)realloc(dp->end, sizeof(int) * dp->rows_m);
        dp->beg_sn = (int*)realloc(dp->beg_sn, sizeof(int) * dp->rows_m);
    }
    size_t used = 0;
#define ROW_ALLOC(r, b, e) do { \
        size_t need = used + (size_t)NPL * ((e) - (b) + 1); \
        if (need > dp->mem_m) { dp->mem_m = need * 2 + (1 << 20); dp->mem = (int32_t*)realloc(dp->mem, sizeof(int32_t) * dp->mem_m); } \
        dp->off[r] = (int64_t)used; dp->beg[r] = (b); dp->end[r] = (e); dp->beg_sn[r] = (b) / pn; used = need; } while (0)
#define PL(r, p) (dp->mem + dp->off[r] + (int64_t)(p) * (dp->end[r] - dp->beg[r] + 1) - dp->beg[r])
    int64_t inband = 0, edge_rows = 0;
    {   /* first row */
        int end0;
        if (wb >= 0) {
            g->mpl[SRC_ID] = g->mpr[SRC_ID] = 0;
            for (i = 0; i < g->node[SRC_ID].out.n; ++i) { int o = g->node[SRC_ID].out.id[i]; g->mpl[o] = g->mpr[o] = 1; }
            int r = qlen - (g->remain[SRC_ID] - g->remain[SINK_ID] - 1);
            end0 = MIN2(qlen, MAX2(g->mpr[SRC_ID], r) + w);
        } else end0 = qlen;
        ROW_ALLOC(0, 0, end0);
        int32_t *H = PL(0, 0);
        if (linear) {
            for (j = 0; j <= end0; ++j) H[j] = local ? 0 : WRAP(-e1 * j);
        } else {
            int32_t *E1 = PL(0, 1), *F1 = PL(0, 2);
            if (local) { for (j = 0; j <= end0; ++j) H[j] = E1[j] = F1[j] = 0; }
            else {
                H[0] = 0; E1[0] = WRAP(-oe1); F1[0] = inf_min;
                for (j = 1; j <= end0; ++j) { E1[j] = inf_min; F1[j] = H[j] = WRAP(-o1 - e1 * j); }
            }
        }
        inband += end0 + 1;
    }
    int32_t best_score = inf_min; int best_i = 0, best_j = 0;
    for (i = 1; i < rows; ++i) {
        int v = g->idx2id[i];
        const elist_t *in = &g->node[v].in;
        int beg, end;
        if (wb < 0) { beg = 0; end = qlen; }
        else {
            int r = qlen - (g->remain[v] - g->remain[SINK_ID] - 1);
            beg = MAX2(0, MIN2(g->mpl[v], r) - w);
            end = MIN2(qlen, MAX2(g->mpr[v], r) + w);
            int beg_sn = beg / pn, min_pre_beg = INT_MAX, min_pre_beg_sn = INT_MAX;
            for (k = 0; k < in->n; ++k) {
                int pi = g->id2idx[in->id[k]];
                if (min_pre_beg > dp->beg[pi]) { min_pre_beg = dp->beg[pi]; min_pre_beg_sn = dp->beg_sn[pi]; }
            }
            if (beg_sn < min_pre_beg_sn) beg = min_pre_beg;
        }
        if (end < beg) end = beg;
        ROW_ALLOC(i, beg, end);
        int32_t *H = PL(i, 0);
        const int *mrow = mat + 5 * g->node[v].base;
        inband += end - beg + 1; edge_rows += (int64_t)in->n * (end - beg + 1);
        if (linear) {
            for (j = beg; j <= end; ++j) H[j] = inf_min;
            for (k = 0; k < in->n; ++k) {
                int pi = g->id2idx[in->id[k]];
                const int32_t *pH = PL(pi, 0);
                int pb = dp->beg[pi], pe = dp->end[pi];
                for (j = beg; j <= end; ++j) {
                    int32_t s = j == 0 ? 0 : mrow[query[j - 1]];
                    int32_t m = (j - 1 >= pb && j - 1 <= pe) ? pH[j - 1] : inf_min;
                    if (local && j == 0) m = 0;  /* :760 `first` = 0 */
                    int32_t d = (j >= pb && j <= pe) ? pH[j] : inf_min;
                    int32_t c = MAX2(WRAP(m + s), WRAP(d - e1));
                    /* the reference only visits vectors that overlap the predecessor's (:764-771); cells it skips keep
                     * inf_min, cells it visits with both operands out of range get inf_min + s or inf_min - e1: junk either way */
                    if (j - 1 > pe + pn || j < pb - pn) continue;
                    if (c > H[j]) H[j] = c;
                }
            }
            for (j = beg + 1; j <= end; ++j) { int32_t f = WRAP(H[j - 1] - e1); if (f > H[j]) H[j] = f; }
            if (local) for (j = beg; j <= end; ++j) if (H[j] < 0) H[j] = 0;
        } else {
            int32_t *E1 = PL(i, 1), *F1 = PL(i, 2);
            for (j = beg; j <= end; ++j) { H[j] = inf_min; E1[j] = inf_min; }
            for (k = 0; k < in->n; ++k) {
                int pi = g->id2idx[in->id[k]];
                const int32_t *pH = PL(pi, 0), *pE1 = PL(pi, 1);
                int pb = dp->beg[pi], pe = dp->end[pi];
                int lo = MAX2(beg, pb + 1), hi = MIN2(end, pe + 1);
                for (j = lo; j <= hi; ++j) if (pH[j - 1] > H[j]) H[j] = pH[j - 1];
                if (local && beg == 0 && 0 > H[0]) H[0] = 0;
                lo = MAX2(beg, pb); hi = MIN2(end, pe);
                for (j = lo; j <= hi; ++j) if (pE1[j] > E1[j]) E1[j] = pE1[j];
            }
            int32_t prevHm = inf_min, f1 = inf_min;
            for (j = beg; j <= end; ++j) {
                int32_t s = j == 0 ? 0 : mrow[query[j - 1]];
                int32_t hm = WRAP(H[j] + s);
                if (j == beg) f1 = (beg % pn == 0) ? WRAP(hm - oe1) : WRAP(inf_min - oe1);  /* :898,:908 */
                else f1 = MAX2(WRAP(prevHm - oe1), WRAP(f1 - e1));
                F1[j] = f1;
                int32_t t = MAX2(hm, E1[j]);
                int32_t h = MAX2(t, f1);
                if (local) h = MAX2(h, 0);
                H[j] = h;
                E1[j] = (h == t) ? MAX2(WRAP(E1[j] - e1), WRAP(h - oe1)) : (local ? 0 : inf_min);  /* :926,:930 */
                prevHm = hm;
            }
        }
        if (local || wb >= 0) {
            int32_t mx = inf_min; int left = -1, right = -1;
            for (j = beg; j <= end; ++j) {
                if (H[j] > mx) { mx = H[j]; left = right = j; }
                else if (H[j] == mx) right = j;
            }
            if (local && mx > best_score) { best_score = mx; best_i = i; best_j = left; }
            if (wb >= 0) {
                const elist_t *out = &g->node[v].out;
                for (k = 0; k < out->n; ++k) {
                    int o = out->id[k];
                    if (right + 1 > g->mpr[o]) g->mpr[o] = right + 1;
                    if (left + 1 < g->mpl[o]) g->mpl[o] = left + 1;
                }
            }
        }
    }
    if (!local) {
        const elist_t *in = &g->node[SINK_ID].in;
        for (k = 0; k < in->n; ++k) {
            int pi = g->id2idx[in->id[k]];
            int e = qlen > dp->end[pi] ? dp->end[pi] : qlen;
            int32_t sc = PL(pi, 0)[e];
            if (sc > best_score) { best_score = sc; best_i = pi; best_j = e; }
        }
    }
    res->best_score = best_score;
    res->inband = inband; res->edge_rows = edge_rows; res->full = (int64_t)rows * (qlen + 1);
    {   /* traceback */
        int cur_op = OP_ALL, hit, id, s;
        i = best_i; j = best_j; id = g->idx2id[i];
        if (best_j < qlen) push_cigar(res, CINS, qlen - best_j, -1, qlen - 1);
        while (i > 0 && j > 0) {
            const int32_t *H = PL(i, 0);
            if (local && H[j] == 0) break;
            const elist_t *in = &g->node[id].in;
            s = mat[5 * g->node[id].base + query[j - 1]]; hit = 0;
            if (linear || (cur_op & OP_M)) {
                for (k = 0; k < in->n; ++k) {
                    int pi = g->id2idx[in->id[k]];
                    if (j - 1 < dp->beg[pi] || j - 1 > dp->end[pi]) continue;
                    if (WRAP(PL(pi, 0)[j - 1] + s) == H[j]) {
                        push_cigar(res, CMATCH, 1, id, j - 1);
                        i = pi; --j; id = g->idx2id[i]; hit = 1; cur_op = OP_ALL;
                        break;
                    }
                }
            }
            if (linear) {
                if (hit == 0) for (k = 0; k < in->n; ++k) {
                    int pi = g->id2idx[in->id[k]];
                    if (j < dp->beg[pi] || j > dp->end[pi]) continue;
                    if (WRAP(PL(pi, 0)[j] - e1) == H[j]) { push_cigar(res, CDEL, 1, id, j - 1); i = pi; id = g->idx2id[i]; hit = 1; break; }
                }
                if (hit == 0) {
                    int32_t hl = (j - 1 >= dp->beg[i]) ? H[j - 1] : inf_min;
                    if (WRAP(hl - e1) == H[j]) { push_cigar(res, CINS, 1, id, j - 1); --j; hit = 1; }
                }
            } else {
                const int32_t *E1 = PL(i, 1), *F1 = PL(i, 2);
                if (hit == 0 && (cur_op & OP_E1)) {
                    for (k = 0; k < in->n; ++k) {
                        int pi = g->id2idx[in->id[k]];
                        if (j < dp->beg[pi] || j > dp->end[pi]) continue;
                        const int32_t *pH = PL(pi, 0), *pE1 = PL(pi, 1);
                        int cond = (cur_op & OP_M) ? (H[j] == pE1[j]) : (E1[j] == WRAP(pE1[j] - e1));
                        if (cond) {
                            if (WRAP(pH[j] - oe1) == pE1[j]) cur_op = OP_M | OP_F; else cur_op = OP_E1;
                            hit = 1; push_cigar(res, CDEL, 1, id, j - 1);
                            i = pi; id = g->idx2id[i];
                            break;
                        }
                    }
                }
                if (hit == 0 && (cur_op & OP_F)) {
                    int32_t hl = (j - 1 >= dp->beg[i]) ? H[j - 1] : inf_min;
                    int32_t f1l = (j - 1 >= dp->beg[i]) ? F1[j - 1] : inf_min;
                    if (!(cur_op & OP_M) || H[j] == F1[j]) {
                        if (WRAP(hl - oe1) == F1[j]) { cur_op = OP_M | OP_E; hit = 1; }
                        else if (WRAP(f1l - e1) == F1[j]) { cur_op = OP_F1; hit = 1; }
                    }
                    if (hit == 1) { push_cigar(res, CINS, 1, id, j - 1); --j; }
                }
            }
            if (hit == 0) { fprintf(stderr, "[poa_oracle] %s backtrack dead end at (%d,%d) cur_op=%d\n", linear ? "lg" : "ag", i, j, cur_op); abort(); }
        }
        if (j > 0) push_cigar(res, CINS, j, -1, j - 1);
        for (k = 0; k < res->n_cigar >> 1; ++k) {
            uint64_t t = res->cigar[k]; res->cigar[k] = res->cigar[res->n_cigar - 1 - k]; res->cigar[res->n_cigar - 1 - k] = t;
        }
    }
#undef ROW_ALLOC
#undef PL
}

/* abpoa_graph.c:688-773 (+ :573-592 for the first sequence); records qpos -> node id in path[] */
static void add_alignment(graph_t *g, int banded, const uint8_t *seq, int w, int seq_l, const aln_t *res, int have_aln, int *path, int *path_len) {
    int i, j;
    *path_len = 0;
    if (g->n == 2) {
        int last = SRC_ID;
        for (i = 0; i < seq_l; ++i) {
            int cur = add_node(g, seq[i]);
            path[i] = cur;
            add_edge(g, last, cur, 0, w);
            last = cur;
        }
        add_edge(g, last, SINK_ID, 0, w);
        *path_len = seq_l;
        topological_sort(g, banded);
        return;
    }
    if (!have_aln || res->n_cigar == 0) return; /* abpoa_graph.c:706-708: read is silently not added */
    int query_id = -1, last_new = 0, last_id = SRC_ID;
    for (i = 0; i < res->n_cigar; ++i) {
        int op = (int)(res->cigar[i] & 0xf);
        if (op == CMATCH) {
            int node_id = (int)((res->cigar[i] >> 34) & 0x3fffffff);
            query_id++;
            if (g->node[node_id].base != seq[query_id]) {
                int aligned_id = get_aligned_id(g, node_id, seq[query_id]);
                if (aligned_id != -1) {
                    add_edge(g, last_id, aligned_id, 1 - last_new, w);
                    last_id = aligned_id; last_new = 0;
                } else {
                    int new_id = add_node(g, seq[query_id]);
                    add_edge(g, last_id, new_id, 0, w);
                    last_id = new_id; last_new = 1;
                    add_aligned(g, node_id, new_id);
                }
            } else {
                add_edge(g, last_id, node_id, 1 - last_new, w);
                last_id = node_id; last_new = 0;
            }
            path[query_id] = last_id;
        } else if (op == CINS) {
            int len = (int)((res->cigar[i] >> 4) & 0x3fffffff);
            query_id += len;
            for (j = len - 1; j >= 0; --j) {
                int new_id = add_node(g, seq[query_id - j]);
                add_edge(g, last_id, new_id, 0, w);
                last_id = new_id; last_new = 1;
                path[query_id - j] = last_id;
            }
        }
    }
    add_edge(g, last_id, SINK_ID, 1 - last_new, w);
    *path_len = seq_l;
    topological_sort(g, banded);
}

/* abpoa_output.c:468-536 + :375-391, n_clu == 1 */
static int heaviest_bundling(graph_t *g, int *cons) {
    int n = g->n, i;
    int *outdeg = (int*)malloc(sizeof(int) * n), *score = (int*)calloc(n, sizeof(int)), *max_out = (int*)malloc(sizeof(int) * n);
    int *q = (int*)malloc(sizeof(int) * (n + 8)); int qh = 0, qt = 0;
    for (i = 0; i < n; ++i) { outdeg[i] = g->node[i].out.n; max_out[i] = -1; }
    q[qt++] = SINK_ID;
    while (qh < qt) {
        int cur = q[qh++];
        if (cur == SINK_ID) { max_out[cur] = -1; score[cur] = 0; }
        else {
            int max_id = -1;
            const elist_t *out = &g->node[cur].out;
            if (cur == SRC_ID) {
                int path_score = -1, path_max_w = -1;
                for (i = 0; i < out->n; ++i) {
                    int o = out->id[i], ow = out->w[i];
                    if (ow > path_max_w || (ow == path_max_w && score[o] > path_score)) { max_id = o; path_score = score[o]; path_max_w = ow; }
                }
                max_out[cur] = max_id;
                break;
            } else {
                int max_w = INT32_MIN;
                for (i = 0; i < out->n; ++i) {
                    int o = out->id[i], ow = out->w[i];
                    if (max_w < ow) { max_w = ow; max_id = o; }
                    else if (max_w == ow && score[max_id] <= score[o]) max_id = o;
                }
                score[cur] = max_w + score[max_id];
                max_out[cur] = max_id;
            }
        }
        for (i = 0; i < g->node[cur].in.n; ++i) { int p = g->node[cur].in.id[i]; if (--outdeg[p] == 0) q[qt++] = p; }
    }
    int len = 0, cur = max_out[SRC_ID];
    while (cur != SINK_ID && cur >= 0) { cons[len++] = cur; cur = max_out[cur]; }
    free(outdeg); free(score); free(max_out); free(q);
    return len;
}

/* abpoa_graph.c:359-410: LIFO walk; an aligned group shares one rank */
static void set_msa_rank(graph_t *g, int *rank) {
    int n = g->n, i, j, msa_rank = 0;
    int *indeg = (int*)malloc(sizeof(int) * n);
    int *st = (int*)malloc(sizeof(int) * (n + 8)); int sp = 0;
    for (i = 0; i < n; ++i) { indeg[i] = g->node[i].in.n; rank[i] = -1; }
    st[sp++] = SRC_ID; rank[SRC_ID] = -1;
    while (sp > 0) {
        int cur = st[--sp];
        if (rank[cur] < 0) {
            rank[cur] = msa_rank;
            for (i = 0; i < g->node[cur].aln_n; ++i) rank[g->node[cur].aln[i]] = msa_rank;
            msa_rank++;
        }
        if (cur == SINK_ID) break;
        for (i = 0; i < g->node[cur].out.n; ++i) {
            int o = g->node[cur].out.id[i];
            if (--indeg[o] == 0) {
                int ok = 1;
                for (j = 0; j < g->node[o].aln_n; ++j) if (indeg[g->node[o].aln[j]] != 0) { ok = 0; break; }
                if (!ok) continue;
                st[sp++] = o; rank[o] = -1;
                for (j = 0; j < g->node[o].aln_n; ++j) { st[sp++] = g->node[o].aln[j]; rank[g->node[o].aln[j]] = -1; }
            }
        }
    }
    free(indeg); free(st);
}

static void graph_free(graph_t *g) {
    int i;
    for (i = 0; i < g->m; ++i) { free(g->node[i].in.id); free(g->node[i].in.w); free(g->node[i].out.id); free(g->node[i].out.w); }
    free(g->node); free(g->idx2id); free(g->id2idx); free(g->mpl); free(g->mpr); free(g->remain);
}

int32_t *oracle_poa_block(const pd_params_t *P, int n_seq, const int32_t *seq_len, const uint8_t *bases,
                          const int32_t *weight, int instrument, int64_t *n_out) {
    (void)instrument;
    *n_out = 0;
    const int gap_mode = P->gap_open1 == 0 ? 2 : (P->gap_open2 == 0 ? 1 : 0); /* abpoa_align.c:87-91: 0 convex, 1 affine, 2 linear */
    int mat[25], i, j, k;
    { /* abpoa_align.c:12-25 */
        int match = P->match < 0 ? -P->match : P->match;
        int mismatch = P->mismatch > 0 ? -P->mismatch : P->mismatch;
        for (i = 0; i < 4; ++i) { for (j = 0; j < 4; ++j) mat[i * 5 + j] = i == j ? match : mismatch; mat[i * 5 + 4] = 0; }
        for (j = 0; j < 5; ++j) mat[20 + j] = 0;
    }
    const int local = P->align_mode == 1;
    const int banded = !local && P->wb >= 0;
    graph_t g; memset(&g, 0, sizeof(g));
    add_node(&g, 0); add_node(&g, 0);
    dpm_t dp; memset(&dp, 0, sizeof(dp));
    int64_t tot_len = 0;
    for (i = 0; i < n_seq; ++i) tot_len += seq_len[i];
    int *paths = (int*)malloc(sizeof(int) * (tot_len + 1));
    int *plen = (int*)calloc(n_seq + 1, sizeof(int));
    int32_t *best = (int32_t*)calloc(n_seq + 1, sizeof(int32_t)), *ncig = (int32_t*)calloc(n_seq + 1, sizeof(int32_t));
    pd_buf_t cig = {0, 0, 0};
    int64_t inband = 0, full = 0, edge_rows = 0, off = 0;
    for (i = 0; i < n_seq; ++i) {
        aln_t res; memset(&res, 0, sizeof(res));
        int have = 0;
        if (g.n > 2) { /* abpoa_align.c:193-198 */
            if (gap_mode == 0) align_sequence(&g, P, mat, bases + off, seq_len[i], &dp, &res);
            else align_sequence_al(&g, P, mat, bases + off, seq_len[i], &dp, &res, gap_mode == 2);
            have = 1;
            inband += res.inband; full += res.full; edge_rows += res.edge_rows;
            best[i] = res.best_score; ncig[i] = res.n_cigar;
            for (k = 0; k < res.n_cigar; ++k) {
                pd_push(&cig, (int32_t)(uint32_t)(res.cigar[k] & 0xffffffffULL));
                pd_push(&cig, (int32_t)(uint32_t)(res.cigar[k] >> 32));
            }
        }
        add_alignment(&g, banded, bases + off, weight[i], seq_len[i], &res, have, paths + off, &plen[i]);
        free(res.cigar);
        off += seq_len[i];
    }
    /* consensus / msa */
    int cons_len = -1, msa_len = -1, msa_rows = 0;
    int *cons = (int*)malloc(sizeof(int) * (g.n + 1));
    uint8_t *msa = NULL;
    if ((P->out_cons) && g.n > 2) cons_len = heaviest_bundling(&g, cons);
    else if (P->out_cons) cons_len = 0;
    if (P->out_msa) {
        if (g.n > 2) {
            int *rank = (int*)malloc(sizeof(int) * g.n);
            set_msa_rank(&g, rank);
            msa_len = rank[SINK_ID] - 1;
            msa_rows = n_seq + (P->out_cons ? 1 : 0);
            if (msa_len <= 0) { msa_len = 0; msa_rows = 0; }
            msa = (uint8_t*)malloc((size_t)msa_rows * msa_len + 1);
            memset(msa, 5, (size_t)msa_rows * msa_len + 1);
            off = 0;
            for (i = 0; i < n_seq; ++i) {
                for (j = 0; j < plen[i]; ++j) { int nd = paths[off + j]; msa[(size_t)i * msa_len + rank[nd] - 1] = g.node[nd].base; }
                off += seq_len[i];
            }
            if (P->out_cons && msa_rows > 0) for (j = 0; j < cons_len; ++j) msa[(size_t)n_seq * msa_len + rank[cons[j]] - 1] = g.node[cons[j]].base;
            free(rank);
        } else { msa_len = 0; msa_rows = 0; }
    }
    /* ---- dump */
    pd_buf_t b = {0, 0, 0};
    for (i = 0; i < PD_HEADER_LEN; ++i) pd_push(&b, 0);
    int64_t n_in = 0, n_oe = 0, n_aln = 0, path_tot = 0;
    for (i = 0; i < g.n; ++i) pd_push(&b, g.node[i].base);
    for (i = 0; i < g.n; ++i) { pd_push(&b, g.node[i].in.n); n_in += g.node[i].in.n; }
    for (i = 0; i < g.n; ++i) for (j = 0; j < g.node[i].in.n; ++j) pd_push(&b, g.node[i].in.id[j]);
    for (i = 0; i < g.n; ++i) for (j = 0; j < g.node[i].in.n; ++j) pd_push(&b, g.node[i].in.w[j]);
    for (i = 0; i < g.n; ++i) { pd_push(&b, g.node[i].out.n); n_oe += g.node[i].out.n; }
    for (i = 0; i < g.n; ++i) for (j = 0; j < g.node[i].out.n; ++j) pd_push(&b, g.node[i].out.id[j]);
    for (i = 0; i < g.n; ++i) for (j = 0; j < g.node[i].out.n; ++j) pd_push(&b, g.node[i].out.w[j]);
    for (i = 0; i < g.n; ++i) { pd_push(&b, g.node[i].aln_n); n_aln += g.node[i].aln_n; }
    for (i = 0; i < g.n; ++i) for (j = 0; j < g.node[i].aln_n; ++j) pd_push(&b, g.node[i].aln[j]);
    for (i = 0; i < n_seq; ++i) { pd_push(&b, plen[i]); path_tot += plen[i]; }
    off = 0;
    for (i = 0; i < n_seq; ++i) { for (j = 0; j < plen[i]; ++j) pd_push(&b, paths[off + j]); off += seq_len[i]; }
    for (i = 0; i < (cons_len > 0 ? cons_len : 0); ++i) pd_push(&b, cons[i]);
    for (int64_t q = 0; q < (int64_t)msa_rows * (msa_len > 0 ? msa_len : 0); ++q) pd_push(&b, msa[q]);
    for (i = 0; i < n_seq; ++i) pd_push(&b, best[i]);
    for (i = 0; i < n_seq; ++i) pd_push(&b, ncig[i]);
    for (int64_t q = 0; q < cig.n; ++q) pd_push(&b, cig.d[q]);
    b.d[PD_MAGIC] = POA_DUMP_MAGIC; b.d[PD_N_NODE] = g.n; b.d[PD_N_SEQ] = n_seq;
    b.d[PD_CONS_LEN] = cons_len; b.d[PD_MSA_LEN] = msa_len; b.d[PD_MSA_ROWS] = msa_rows;
    b.d[PD_N_IN_TOT] = (int32_t)n_in; b.d[PD_N_OUT_TOT] = (int32_t)n_oe; b.d[PD_N_ALN_TOT] = (int32_t)n_aln;
    b.d[PD_PATH_TOT] = (int32_t)path_tot; b.d[PD_CIGAR_TOT] = (int32_t)(cig.n / 2);
    pd_set64(b.d, PD_INBAND_LO, inband); pd_set64(b.d, PD_FULL_LO, full); pd_set64(b.d, PD_EDGE_ROWS_LO, edge_rows);
    free(paths); free(plen); free(best); free(ncig); free(cig.d); free(cons); free(msa);
    free(dp.mem); free(dp.off); free(dp.beg); free(dp.end); free(dp.beg_sn);
    graph_free(&g);
    *n_out = b.n;
    return b.d;
}
