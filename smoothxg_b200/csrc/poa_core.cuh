// poa_core.cuh -- device code of the B200-native POA engine (one CUDA block per POA block).
//
// What it computes is abPOA v1.5.4's partial order alignment (convex, affine or linear gaps) exactly as
// smoothxg drives it (reference citations relative to /root/reference):
//   per-block driver loop          deps/abPOA/src/abpoa_align.c:304-344        -> poa_block()
//   int16/int32 choice, inf_min    deps/abPOA/src/abpoa_align_simd.c:1286-1302 -> poa_block()
//   first row / row recurrence     deps/abPOA/src/abpoa_align_simd.c:617-688, :935-1074 (cg), :817-933 (ag), :727-815 (lg)
//                                  -> fill<NW,S,MODE>() here, fill_p16<>() in poa_fill16.cuh (packed 16-bit, convex)
//   adaptive band                  deps/abPOA/src/abpoa_align.h:34-35, abpoa_align_simd.c:1107-1130
//   best cell                      deps/abPOA/src/abpoa_align_simd.c:1092-1105, :1208-1210
//   backtrack                      deps/abPOA/src/abpoa_align_simd.c:309-458 (cg), :196-307 (ag), :116-194 (lg) -> backtrack<>(), bt_step<>()
//   graph fusion                   deps/abPOA/src/abpoa_graph.c:688-773, :480-556, :573-592 -> fuse_par()
//   topological sort               deps/abPOA/src/abpoa_graph.c:322-357 (:221-266, :192-219, :268-309) -> toposort() (local mode),
//                                  toposort_incr() (global mode: incremental order, pointer-jumping `remain`)
//   heaviest-bundle consensus      deps/abPOA/src/abpoa_output.c:468-536, :375-391
//   MSA rank                       deps/abPOA/src/abpoa_graph.c:359-419
//
// How it computes it is not the reference's: a DP row is evaluated by all lanes at once in absolute column
// coordinates (so predecessor rows line up without shifts), the horizontal gap recurrences F1/F2 are max-plus
// prefix scans (G[j] = F[j] + e*j turns F[j] = max(F[j-1]-e, H[j-1]-oe) into a running maximum) done with warp
// shuffles, the row maximum / arg-max for the adaptive band is a redux.sync reduction, rows are stored band-only,
// traceback commits diagonal runs 32 steps at a time, fusion and the order update are data-parallel, and the
// whole per-block loop (align, traceback, fusion, order, consensus, MSA) stays on the device.
//
// The file also compiles as plain C++ with -DPOA_HOST_EMU: one emulated thread (warp size 1), or with
// -DPOA_EMU_LANES=32 thirty-two lock-step lanes run as fibers by the test harness (tests/emu/emu_poa.cpp), so
// the warp-level logic is checked against the golden vectors on a machine without a GPU.  Those builds are test
// infrastructure: never built by __graft_entry__.build(), never shipped, never reachable from the C ABI.
#pragma once
#include <stdint.h>
#include <limits.h>

#ifdef POA_HOST_EMU
#include <vector_types.h>
#include <string.h>
#define POA_D static inline
#define POA_DN static
#define POA_SM static inline
#ifndef POA_EMU_LANES
#define POA_EMU_LANES 1
#endif
#define POA_WARP POA_EMU_LANES
#if POA_EMU_LANES == 1
static inline int poa_tid() { return 0; }
static inline void poa_sync_block() {}
static inline void poa_sync_warp() {}
static inline int poa_shfl_up(int v, int) { return v; }
static inline int poa_shfl(int v, int) { return v; }
static inline int poa_redux_max(int v) { return v; }
static inline int poa_redux_min(int v) { return v; }
static inline unsigned poa_ballot(int p) { return p ? 1u : 0u; }
#else
// Lock-step lanes emulated as fibers by the test harness (tests/emu/emu_poa.cpp): POA_EMU_NW warps of 32 lanes.
// Every warp collective is an exchange among the 32 fibers of one warp, a block barrier a rendezvous of all of
// them, so the warp- and block-level logic (shuffle scans, reductions, lane-striped layouts, cross-warp
// exchanges through shared memory) runs on a machine without a GPU.
namespace poa_emu { int lane(); void xchg(int v, int *all); void sync_all(); }
static inline int poa_tid() { return poa_emu::lane(); }
static inline void poa_sync_warp() { int a[32]; poa_emu::xchg(0, a); }
static inline void poa_sync_block() { poa_emu::sync_all(); }
static inline int poa_shfl_up(int v, int d) { int a[32]; poa_emu::xchg(v, a); int l = poa_emu::lane() & 31; return l >= d ? a[l - d] : v; }
static inline int poa_shfl(int v, int l) { int a[32]; poa_emu::xchg(v, a); return a[l & 31]; }
static inline int poa_redux_max(int v) { int a[32]; poa_emu::xchg(v, a); int m = a[0]; for (int i = 1; i < 32; ++i) m = a[i] > m ? a[i] : m; return m; }
static inline int poa_redux_min(int v) { int a[32]; poa_emu::xchg(v, a); int m = a[0]; for (int i = 1; i < 32; ++i) m = a[i] < m ? a[i] : m; return m; }
static inline unsigned poa_ballot(int p) { int a[32]; poa_emu::xchg(p ? 1 : 0, a); unsigned m = 0; for (int i = 0; i < 32; ++i) if (a[i]) m |= 1u << i; return m; }
#endif
static inline long long poa_clock() { return 0; }
static inline unsigned long long poa_atomic_add(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
static inline int poa_atomic_add(int *p, int v) { int o = *p; *p += v; return o; }
static inline int4 poa_make_int4(int x, int y, int z, int w) { int4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline void poa_red_max(int *p, int v) { if (v > *p) *p = v; }
static inline void poa_red_min(int *p, int v) { if (v < *p) *p = v; }
#else
#define POA_D __device__ __forceinline__
#define POA_DN __device__ __noinline__
#define POA_SM __device__ __forceinline__ static
#define POA_WARP 32
POA_D int poa_tid() { return threadIdx.x; }
POA_D void poa_sync_block() { __syncthreads(); }
POA_D void poa_sync_warp() { __syncwarp(); }
POA_D int poa_shfl_up(int v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
POA_D int poa_shfl(int v, int l) { return __shfl_sync(0xffffffffu, v, l); }
POA_D int poa_redux_max(int v) { return __reduce_max_sync(0xffffffffu, v); }
POA_D int poa_redux_min(int v) { return __reduce_min_sync(0xffffffffu, v); }
POA_D unsigned poa_ballot(int p) { return __ballot_sync(0xffffffffu, p); }
POA_D long long poa_clock() { return clock64(); }
POA_D unsigned long long poa_atomic_add(unsigned long long *p, unsigned long long v) { return atomicAdd(p, v); }
POA_D int poa_atomic_add(int *p, int v) { return atomicAdd(p, v); }
POA_D int4 poa_make_int4(int x, int y, int z, int w) { return make_int4(x, y, z, w); }
POA_D void poa_red_max(int *p, int v) { atomicMax(p, v); }
POA_D void poa_red_min(int *p, int v) { atomicMin(p, v); }
#endif

#ifdef POA_HOST_EMU
#define POA_HD static inline
#else
#define POA_HD __host__ __device__ __forceinline__
#endif

namespace poa {

constexpr int SRC_ID = 0, SINK_ID = 1;
constexpr int HDR_WORDS = 20;  // per-block result header, int32 words
constexpr int NEG_INF32 = INT_MIN;
constexpr int MAX_WARPS = 8;

// per-block status (keep in sync with include/poa_b200.h)
enum { ST_OK = 0, ST_ESLAB = 1, ST_EARENA = 2, ST_EINTERNAL = 3, ST_EUNSUP = 4 };
// result header slots
enum { H_STATUS = 0, H_N_NODE, H_N_SEQ, H_CONS_LEN, H_MSA_LEN, H_MSA_ROWS, H_IN_TOT, H_OUT_TOT, H_ALN_TOT,
       H_PATH_TOT, H_CIG_TOT, H_OFF_LO, H_OFF_HI, H_INBAND_LO, H_INBAND_HI, H_EDGE_LO, H_EDGE_HI,
       H_BODY_WORDS /* int32 words of the block body */, H_FORMAT /* WIRE_* */, H_RUN_TOT /* path runs, consensus included */, H_WORDS };
static_assert(H_WORDS == HDR_WORDS, "header slots");

// ------------------------------------------------------------------------------------------------
// Result body of one block in the arena ("wire format"): what travels device -> host and GPU -> GPU.
// A POA result is mostly small numbers and mostly consecutive node ids, so it is stored narrow:
//   * bases and aligned-group sizes as bytes; degrees, node ids and edge weights as 16-bit values whenever the block allows
//     it (fewer than 65 536 nodes, total sequence weight and every sequence length below 65 536 -- WIRE_NARROW), else 32-bit;
//   * the per-read node paths and the consensus path run-length coded: a read's path through the graph is a chain of runs
//     of consecutive node ids (the nodes one earlier read created in one piece), so (first node id, position in the read) per
//     run replaces one id per base -- 32 x 2 000 ids become a few hundred runs.
// 32 x 2 kb blocks shrink from 386 KB to about 60 KB: that is the D2H copy inside every end-to-end call and the payload of the
// multi-GPU gather.  wire_layout() gives the word offset of every section from the header's totals; the kernel writes through
// it (poa_block) and the host decodes through it (poa_wire.hpp) into the flat int32 arrays of poa_b200_block_view_t.
// ------------------------------------------------------------------------------------------------
enum { WIRE_NARROW = 1, WIRE_WIDE = 3 };
struct WireLayout {
    long long o_base, o_aln_n, o_in_n, o_out_n, o_in_id, o_in_w, o_out_id, o_out_w, o_aln_id;  // word offsets from the body start
    long long o_plen, o_best, o_ncig, o_nrun, o_runs, o_cig, o_msa, words;
};
POA_HD void wire_layout(WireLayout &W, long long n, long long n_seq, long long in_tot, long long out_tot, long long aln_tot, long long run_tot,
                       long long cig_tot, long long msa_bytes, int format) {
    const long long ew = format == WIRE_WIDE ? 4 : 2;
    long long o = 0;
    W.o_base = o; o += (n + 3) >> 2;
    W.o_aln_n = o; o += (n + 3) >> 2;
    W.o_in_n = o; o += (n * ew + 3) >> 2;
    W.o_out_n = o; o += (n * ew + 3) >> 2;
    W.o_in_id = o; o += (in_tot * ew + 3) >> 2;
    W.o_in_w = o; o += (in_tot * ew + 3) >> 2;
    W.o_out_id = o; o += (out_tot * ew + 3) >> 2;
    W.o_out_w = o; o += (out_tot * ew + 3) >> 2;
    W.o_aln_id = o; o += (aln_tot * ew + 3) >> 2;
    W.o_plen = o; o += n_seq;
    W.o_best = o; o += n_seq;
    W.o_ncig = o; o += n_seq;
    W.o_nrun = o; o += n_seq + 1;            // runs per read, then of the consensus path
    W.o_runs = o; o += (run_tot * 2 * ew + 3) >> 2;  // (first node id, position) per run
    W.o_cig = o; o += 2 * cig_tot;
    W.o_msa = o; o += (msa_bytes + 3) >> 2;
    W.words = o;
}
// phase cycle counters
enum { PH_ROWS = 0, PH_FILL, PH_BT, PH_FUSE, PH_TOPO, PH_FINAL, PH_TOTAL, PH_SPARE, PH_N };

// cigar ops (deps/abPOA/include/abpoa.h:18-24)
constexpr int CMATCH = 0, CINS = 1, CDEL = 2;
// backtrack state bits (deps/abPOA/src/abpoa_align.h:20-27)
constexpr int OP_M = 0x1, OP_E1 = 0x2, OP_E2 = 0x4, OP_E = 0x6, OP_F1 = 0x8, OP_F2 = 0x10, OP_F = 0x18, OP_ALL = 0x1f;

struct DevParams {
    int mat[25];
    int o1, e1, o2, e2, oe1, oe2;
    int match, min_mis;
    int local, wb;
    float wf;
    int out_cons, out_msa;
    int pn16, pn32;  // lane counts of the reference build whose band-start rule we reproduce (AVX-512BW: 32/16)
    int emit_cigar;
    int p16_ok;  // scoring parameters allow the packed 16-bit fill (poa_fill16.cuh)
    int p16_default;  // ... and are smoothxg's defaults 1,4,6,2,26,1: fill_p16<.., PRESET = true> has them as immediates
    int gap_mode;  // 0 convex, 1 affine, 2 linear (abpoa_set_gap_mode, abpoa_align.c:87-91)
    // Packed (two int16 per word) constants of the 16-bit fill, prepared on the host (build_params): as fields of a
    // __grid_constant__ parameter they are constant-bank operands of the packed-integer instructions instead of a dozen
    // loop-invariant vector registers (the fill runs at 128 registers per thread with nothing to spare).
    unsigned pk_inf, pk_negl, pk_noe1, pk_noe2, pk_ne1, pk_ne2, pk_ne1_2, pk_ne1_3, pk_ne2_2, pk_ne2_3, pk_ncw1, pk_ncw2;
};

// ------------------------------------------------------------------------------------------------
// When 16-bit cells are enough although abPOA itself would switch to 32 bits.
// abPOA picks int32 as soon as max(qlen * match, max(qlen, graph rows) * e1 + o1) could leave the int16 range
// (abpoa_align_simd.c:1293-1302) -- a bound on a gap as long as the GRAPH HAS ROWS, which a deep block passes quickly
// (256 x 8 kb: 28 592 rows) although no cell of a global alignment can be that far down: every node lies on the path of some
// sequence of the block, so it is at most Lmax = max sequence length steps from the source, and cell (row, j) is reachable by
// j inserted bases plus at most Lmax deleted nodes.  Hence every REAL cell (H, and E/F which are >= H - oe) is at least
//   real_min = -(gap(qlen) + gap(Lmax)) - max(oe1, oe2),     gap(k) = min(o1 + e1 k, o2 + e2 k)   (o2 = 0: o1 + e1 k),
// and at most qlen * match.  Cells no real path reaches hold inf_min16-derived junk, at most inf_min16 + qlen * match (a junk
// value grows by at most `match` per row).  If junk_max stays below real_min, junk never wins a max against a real value, so the
// 16-bit fill computes every real cell exactly as abPOA's 32-bit kernel does and the traceback, which only ever follows
// equalities between real cells, takes the same steps.  (What does differ between the two kernels, the vector width of the
// band-start rounding rule, is passed in: pn32 instead of pn16.)  Global mode only; local alignments keep the generic path.
// ------------------------------------------------------------------------------------------------
POA_HD bool p16_safe_for_long_graph(const DevParams &P, long long qlen, long long lmax) {
    if (P.local || P.gap_mode != 0) return false;
    const long long g1q = P.o1 + (long long)P.e1 * qlen, g2q = P.o2 + (long long)P.e2 * qlen;
    const long long g1l = P.o1 + (long long)P.e1 * lmax, g2l = P.o2 + (long long)P.e2 * lmax;
    const long long gq = g1q < g2q ? g1q : g2q, gl = g1l < g2l ? g1l : g2l;
    const long long oe = P.oe1 > P.oe2 ? P.oe1 : P.oe2, emax = P.e1 > P.e2 ? P.e1 : P.e2;
    const long long real_min = -(gq + gl) - oe - P.min_mis;
    long long a = -32768LL + P.min_mis, b = -32768LL + P.oe1, c = -32768LL + P.oe2;
    const long long inf16 = (a > b ? (a > c ? a : c) : (b > c ? b : c)) + 512 * emax;  // inf_min_of<short>()
    const long long junk_max = inf16 + qlen * P.match + 2 * oe;
    return junk_max + 1024 < real_min;
}

// Device-resident batch input (flat, same arrays as the C ABI takes).
struct DevBatch {
    const long long *block_seq_off;
    const int *seq_len;
    const long long *seq_off;
    const uint8_t *bases;
    const int *weight;
    const int *order;  // processing order: block ids, most expensive first
    int n_order;
};

// Byte offsets of the per-CTA workspace arrays (identical for all CTAs of a launch).
struct WsLayout {
    long long stride;  // bytes per CTA
    long long o_base, o_aln_n, o_aln, o_in_off, o_in_n, o_out_off, o_out_n, o_pool_id, o_pool_w, o_pool_row;
    long long o_idx2id, o_id2idx, o_remain, o_tmp0, o_tmp1, o_tmp2, o_tmp3;
    long long o_rowinfo, o_rowmeta, o_rbase, o_rr, o_mplr, o_mprr, o_pred4;
    long long o_cig, o_path, o_best, o_ncig, o_plen, o_nrun, o_qp, o_slab;
    long long slab_bytes;
    unsigned slab_planes;  // 512-byte chunk-planes the slab holds, capped at 0xfff00000 (packed 16-bit fill: 32-bit bookkeeping)
    int nmax;      // node capacity
    int pool_cap;  // edge pool capacity (entries)
    int cig_cap;   // cigar words capacity per block (all sequences when emit_cigar, else one alignment)
};

struct DevOut {
    int *hdr;                       // [n_blocks][HDR_WORDS]
    int *arena;                     // result bodies, int32 words
    unsigned long long *arena_used; // bump cursor (words)
    unsigned long long arena_cap;   // words
    unsigned long long *phase;      // [PH_N] summed cycles
    int *counter;                   // next entry of DevBatch::order to take
};

struct Ws {
    uint8_t *base, *aln_n, *rbase;
    int *aln, *in_off, *in_n, *out_off, *out_n, *pool_id, *pool_w, *pool_row;
    int *idx2id, *id2idx, *remain, *tmp0, *tmp1, *tmp2, *tmp3;
    int4 *rowinfo, *rowmeta;
    int4 *pred4;  // build_rows(): rows of a row's first four predecessors in in_id order, -1 where there is none
    int *rr, *mplr, *mprr;
    unsigned long long *cig;
    int *path, *best, *ncig, *plen;
    int *nrun;   // [2 * (max_seq + 2)]: runs per read path / their exclusive prefix (output stage)
    char *qp;    // query profile of the alignment in flight, chunked layout (poa_fill16.cuh)
    char *slab;
};

struct Shared {
    Ws ws;
    int blk;
    int n_node;
    int nmax, pool_cap;
    int pool_used;
    int err;
    int n_cigar;       // cigar words of the current alignment
    int cig_base;      // where the current alignment's cigar starts in ws.cig
    int best_score, best_i, best_j;
    long long inband, edge_rows;
    int scan_x[2][2][MAX_WARPS];
    int red_x[2][3][MAX_WARPS];
    int bcast[4];
    char *ring;  // previous-row cache of the packed 16-bit fill (dynamic shared memory, poa_fill16.cuh)
    int ring_bytes;
};

#ifdef POA_HOST_EMU
static inline int p_ctz32(unsigned x) { return __builtin_ctz(x); }
static inline int poa_popc(unsigned x) { return __builtin_popcount(x); }
#else
POA_D int p_ctz32(unsigned x) { return __ffs((int)x) - 1; }
POA_D int poa_popc(unsigned x) { return __popc(x); }
#endif
POA_D int imax(int a, int b) { return a > b ? a : b; }
POA_D int imin(int a, int b) { return a < b ? a : b; }

// ------------------------------------------------------------------------------------------------
// graph primitives (single thread)
// ------------------------------------------------------------------------------------------------
// Edge lists live in one pool.  A node's in-list starts at 4*id and its out-list at 4*id+2 (two
// entries each); a list that outgrows its slot moves to a fresh slot of twice the size taken from the
// growth region behind 4*nmax.  The pool is sized for the worst case so it cannot run out.
POA_D void edge_push(Shared &sh, int *off_arr, int *n_arr, int v, int id, int wt) {
    Ws &w = sh.ws;
    int n = n_arr[v], off = off_arr[v];
    if (n >= 2 && (n & (n - 1)) == 0) {
        int noff = sh.pool_used;
        if (noff + 2 * n > sh.pool_cap) { sh.err = ST_ESLAB; return; }
        sh.pool_used += 2 * n;
        for (int t = 0; t < n; ++t) { w.pool_id[noff + t] = w.pool_id[off + t]; w.pool_w[noff + t] = w.pool_w[off + t]; }
        off_arr[v] = noff; off = noff;
    }
    w.pool_id[off + n] = id; w.pool_w[off + n] = wt; n_arr[v] = n + 1;
}

POA_D int add_node(Shared &sh, int base) {  // abpoa_graph.c:471-478
    Ws &w = sh.ws;
    int v = sh.n_node++;
    w.base[v] = (uint8_t)base; w.aln_n[v] = 0;
    w.in_n[v] = 0; w.out_n[v] = 0; w.in_off[v] = 4 * v; w.out_off[v] = 4 * v + 2;
    return v;
}

POA_D void add_edge(Shared &sh, int from, int to, int check_edge, int wt) {  // abpoa_graph.c:480-556
    Ws &w = sh.ws;
    int exist = 0;
    if (check_edge) {
        int n = w.in_n[to], off = w.in_off[to];
        for (int i = 0; i < n; ++i) if (w.pool_id[off + i] == from) { w.pool_w[off + i] += wt; break; }
        n = w.out_n[from]; off = w.out_off[from];
        for (int i = 0; i < n; ++i) if (w.pool_id[off + i] == to) { w.pool_w[off + i] += wt; exist = 1; break; }
    }
    if (!exist) {
        edge_push(sh, w.in_off, w.in_n, to, from, wt);
        edge_push(sh, w.out_off, w.out_n, from, to, wt);
    }
}

POA_D void add_aligned(Ws &w, int node_id, int new_id) {  // abpoa_graph.c:455-463
    int n = w.aln_n[node_id];
    for (int i = 0; i < n; ++i) {
        int a = w.aln[4 * node_id + i];
        w.aln[4 * a + w.aln_n[a]++] = new_id;
        w.aln[4 * new_id + w.aln_n[new_id]++] = a;
    }
    w.aln[4 * node_id + w.aln_n[node_id]++] = new_id;
    w.aln[4 * new_id + w.aln_n[new_id]++] = node_id;
}

POA_D int get_aligned_id(const Ws &w, int node_id, int base) {  // abpoa_graph.c:439-448
    int n = w.aln_n[node_id];
    for (int i = 0; i < n; ++i) {
        int a = w.aln[4 * node_id + i];
        if (w.base[a] == base) return a;
    }
    return -1;
}

// ------------------------------------------------------------------------------------------------
// block-wide helpers
// ------------------------------------------------------------------------------------------------
template <int NW>
POA_D void sync_block() {
    if (NW == 1) poa_sync_warp(); else poa_sync_block();
}

// ------------------------------------------------------------------------------------------------
// topological sort (abpoa_graph.c:322-357)
// ------------------------------------------------------------------------------------------------
template <int NW>
POA_DN void toposort(Shared &sh, int banded) {
    constexpr int NT = NW * POA_WARP;
    Ws &w = sh.ws;
    const int tid = poa_tid();
    const int n = sh.n_node;
    // in-degree copy
    for (int v = tid; v < n; v += NT) w.tmp0[v] = w.in_n[v];
    sync_block<NW>();
    if (tid == 0) {  // abpoa_graph.c:221-266; the FIFO queue *is* index_to_node_id
        int qh = 0, qt = 0;
        w.idx2id[qt++] = SRC_ID;
        while (qh < qt) {
            int cur = w.idx2id[qh];
            w.id2idx[cur] = qh++;
            if (cur == SINK_ID) break;
            int on = w.out_n[cur], ooff = w.out_off[cur];
            for (int i = 0; i < on; ++i) {
                int o = w.pool_id[ooff + i];
                if (--w.tmp0[o] == 0) {
                    int an = w.aln_n[o], ok = 1;
                    for (int j = 0; j < an; ++j) if (w.tmp0[w.aln[4 * o + j]] != 0) { ok = 0; break; }
                    if (!ok) continue;
                    w.idx2id[qt++] = o;
                    for (int j = 0; j < an; ++j) w.idx2id[qt++] = w.aln[4 * o + j];
                }
            }
        }
    }
    sync_block<NW>();
    // abpoa_graph.c:192-219: exchange sort by weight, strict <, not stable; one node per thread
    for (int v = tid; v < n; v += NT) {
        for (int side = 0; side < 2; ++side) {
            int cnt = side ? w.out_n[v] : w.in_n[v];
            int off = side ? w.out_off[v] : w.in_off[v];
            for (int j = 0; j < cnt - 1; ++j)
                for (int k = j + 1; k < cnt; ++k)
                    if (w.pool_w[off + j] < w.pool_w[off + k]) {
                        int t = w.pool_id[off + j]; w.pool_id[off + j] = w.pool_id[off + k]; w.pool_id[off + k] = t;
                        t = w.pool_w[off + j]; w.pool_w[off + j] = w.pool_w[off + k]; w.pool_w[off + k] = t;
                    }
        }
        if (banded) { w.tmp0[v] = w.out_n[v]; w.remain[v] = 0; }
    }
    sync_block<NW>();
    if (banded && tid == 0) {  // abpoa_graph.c:268-309
        int qh = 0, qt = 0;
        int *q = w.tmp1;
        q[qt++] = SINK_ID; w.remain[SINK_ID] = -1;
        while (qh < qt) {
            int cur = q[qh++];
            if (cur != SINK_ID) {
                int max_w = -1, max_id = -1;
                int on = w.out_n[cur], ooff = w.out_off[cur];
                for (int i = 0; i < on; ++i)
                    if (w.pool_w[ooff + i] > max_w) { max_w = w.pool_w[ooff + i]; max_id = w.pool_id[ooff + i]; }
                w.remain[cur] = w.remain[max_id] + 1;
            }
            if (cur == SRC_ID) break;
            int in = w.in_n[cur], ioff = w.in_off[cur];
            for (int i = 0; i < in; ++i) {
                int p = w.pool_id[ioff + i];
                if (--w.tmp0[p] == 0) q[qt++] = p;
            }
        }
    }
    sync_block<NW>();
}

// ------------------------------------------------------------------------------------------------
// Global alignment: row order maintained incrementally instead of re-running the BFS.
//
// abpoa_topological_sort (abpoa_graph.c:322-357) recomputes a BFS order of the whole graph after every
// read; a single-thread pointer chase that is pure latency on a GPU.  In global mode the DP result does
// not depend on WHICH topological order the rows are evaluated in: band limits are min/max
// accumulations over predecessors (abpoa_align_simd.c:1121-1130), `remain` is a property of the graph,
// the best cell and the traceback walk edge lists in in_id order, never row order.  (Local mode breaks
// score ties by BFS index, abpoa_align_simd.c:1208-1210, so it keeps the exact BFS.)  So the order is kept
// as an invariant: topological, aligned groups contiguous.  A fused read visits old nodes in increasing
// order; a node it creates is placed right behind the aligned group of the old node it follows (insertion)
// or is aligned with (mismatch) -- fuse_par() records that old node in anc[].  New positions are old position +
// number of insertions in front: one histogram, one block-wide prefix sum, no serial walk.
// `remain` (abpoa_graph.c:268-309: edges to the sink along heaviest out-edges) is pointer jumping.
// ------------------------------------------------------------------------------------------------
template <int NW>
POA_D int block_any(Shared &sh, int pred) {
    unsigned b = poa_ballot(pred);
    if (NW == 1) return b != 0;
    const int tid = poa_tid();
    sync_block<NW>();
    if (tid == 0) sh.bcast[0] = 0;
    sync_block<NW>();
    if (b != 0 && (tid % POA_WARP) == 0) poa_atomic_add(&sh.bcast[0], 1);
    sync_block<NW>();
    const int r = sh.bcast[0];
    sync_block<NW>();
    return r != 0;
}

template <int NW> POA_D int block_excl_scan(Shared &sh, const int *src, int *dst, int n, int *scratch);

template <int NW>
POA_DN void toposort_incr(Shared &sh, int banded, int n_old) {
    constexpr int NT = NW * POA_WARP;
    Ws &w = sh.ws;
    const int tid = poa_tid();
    const int n = sh.n_node, n_new = n - n_old;
    // abpoa_graph.c:192-219: exchange sort by weight, strict <, not stable; one node per thread, degrees of four
    // nodes fetched together (most lists have one entry and need nothing)
    for (int v0 = tid; v0 < n; v0 += 4 * NT) {
        int cn[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int v = v0 + u * NT;
            cn[u][0] = v < n ? w.in_n[v] : 0; cn[u][1] = v < n ? w.out_n[v] : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int v = v0 + u * NT;
            for (int side = 0; side < 2; ++side) {
                const int cnt = cn[u][side];
                if (cnt < 2) continue;
                const int off = side ? w.out_off[v] : w.in_off[v];
                for (int j = 0; j < cnt - 1; ++j)
                    for (int k = j + 1; k < cnt; ++k)
                        if (w.pool_w[off + j] < w.pool_w[off + k]) {
                            int t = w.pool_id[off + j]; w.pool_id[off + j] = w.pool_id[off + k]; w.pool_id[off + k] = t;
                            t = w.pool_w[off + j]; w.pool_w[off + j] = w.pool_w[off + k]; w.pool_w[off + k] = t;
                        }
            }
        }
    }
    if (n_new > 0) {
        int *hist = w.tmp0, *anc = w.tmp2, *apos = w.tmp3;
        for (int i = tid; i < n_old; i += NT) hist[i] = 0;
        sync_block<NW>();
        for (int k = tid; k < n_new; k += NT) {  // position of the last member of the followed group, old order
            const int x = anc[k];
            int ge = w.id2idx[x];
            const int an = w.aln_n[x];
            for (int j = 0; j < an; ++j) { const int m = w.aln[4 * x + j]; if (m < n_old) ge = imax(ge, w.id2idx[m]); }
            apos[k] = ge;
            poa_atomic_add(&hist[ge], 1);
        }
        sync_block<NW>();
        block_excl_scan<NW>(sh, hist, hist, n_old, w.tmp1);  // hist[i] = nodes inserted in front of old position i
        for (int i = tid; i < n_old; i += NT) w.id2idx[w.idx2id[i]] = i + hist[i];
        for (int k = tid; k < n_new; k += NT) w.id2idx[n_old + k] = apos[k] + 1 + k;  // creation order = path order
        sync_block<NW>();
        for (int v = tid; v < n; v += NT) w.idx2id[w.id2idx[v]] = v;
    }
    sync_block<NW>();
    if (banded) {
        // remain[v] = remain[heaviest out-neighbour] + 1, remain[sink] = -1 (abpoa_graph.c:263-283).  Targets first, by
        // all threads; then warp 0 sweeps the order from the sink down, one warp-wide group of positions at a time:
        // links inside a group are resolved by pointer jumping over shuffles, links to earlier groups read the
        // finished values, kept by position in shared memory while the graph fits (the fill's ring is idle here).
        int *tgt = w.tmp0;
        for (int v0 = tid; v0 < n; v0 += 4 * NT) {
            int oo[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int v = v0 + u * NT; oo[u] = (v < n && v != SINK_ID) ? w.out_off[v] : -1; }
#pragma unroll
            for (int u = 0; u < 4; ++u) if (oo[u] >= 0) oo[u] = w.pool_id[oo[u]];  // heaviest out-edge: first after the weight sort
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int v = v0 + u * NT; if (v < n) tgt[v] = oo[u]; }
        }
        sync_block<NW>();
        if (tid < POA_WARP) {
            const int lane = tid;
            int *rem = (sh.ring != nullptr && (long long)n * 4 <= sh.ring_bytes) ? (int *)sh.ring : w.tmp1;
            for (int hi = n - 1; hi >= 0; hi -= 4 * POA_WARP) {
                int vv[4], ti[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int x = hi - u * POA_WARP - lane; vv[u] = x >= 0 ? w.idx2id[x] : -1; }
#pragma unroll
                for (int u = 0; u < 4; ++u) ti[u] = vv[u] >= 0 ? tgt[vv[u]] : -1;
#pragma unroll
                for (int u = 0; u < 4; ++u) ti[u] = ti[u] >= 0 ? w.id2idx[ti[u]] : -1;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int top = hi - u * POA_WARP, x = top - lane;
                    if (top < 0) break;
                    int acc, ptr = -1;
                    if (ti[u] < 0) acc = -1;                                   // the sink (or a lane past the source end)
                    else if (ti[u] <= top) { acc = 1; ptr = top - ti[u]; }     // inside this group: a lower lane
                    else acc = rem[ti[u]] + 1;
                    for (int r = 1; r < POA_WARP; r <<= 1) {
                        const int src = ptr >= 0 ? ptr : lane;
                        const int a = poa_shfl(acc, src), pp = poa_shfl(ptr, src);
                        if (ptr >= 0) { acc += a; ptr = pp; }
                    }
                    if (x >= 0) { rem[x] = acc; w.remain[vv[u]] = acc; }
                    poa_sync_warp();
                }
            }
        }
    }
    sync_block<NW>();
}

// ------------------------------------------------------------------------------------------------
// first sequence (abpoa_graph.c:573-592): a chain src -> bases -> sink, built by all threads
// ------------------------------------------------------------------------------------------------
template <int NW>
POA_DN void add_first_sequence(Shared &sh, const uint8_t *q, int qlen, int wt, int *path) {
    constexpr int NT = NW * POA_WARP;
    Ws &w = sh.ws;
    const int tid = poa_tid();
    // graph holds only src and sink here; sequential semantics: nodes 2..qlen+1 in order.
    // (src/sink may already carry src->sink edges from earlier empty sequences.)
    const int n0 = sh.n_node;
    if (n0 + qlen > sh.nmax) { sync_block<NW>(); if (tid == 0) sh.err = ST_ESLAB; sync_block<NW>(); return; }
    for (int t = tid; t < qlen; t += NT) {
        int v = n0 + t;
        w.base[v] = q[t]; w.aln_n[v] = 0;
        w.in_off[v] = 4 * v; w.out_off[v] = 4 * v + 2;
        w.in_n[v] = 1; w.out_n[v] = 1;
        w.pool_id[4 * v] = (t == 0) ? SRC_ID : v - 1; w.pool_w[4 * v] = wt;
        w.pool_id[4 * v + 2] = (t == qlen - 1) ? SINK_ID : v + 1; w.pool_w[4 * v + 2] = wt;
        path[t] = v;
        w.tmp2[t] = SRC_ID;  // toposort_incr(): the chain goes right behind the source
    }
    sync_block<NW>();
    if (tid == 0) {
        sh.n_node = n0 + qlen;
        if (qlen > 0) {
            edge_push(sh, w.out_off, w.out_n, SRC_ID, n0, wt);
            edge_push(sh, w.in_off, w.in_n, SINK_ID, n0 + qlen - 1, wt);
        } else {
            add_edge(sh, SRC_ID, SINK_ID, 0, wt);
        }
    }
    sync_block<NW>();
}

// ------------------------------------------------------------------------------------------------
// per-alignment row tables: everything the fill loop needs, indexed by row (= BFS index)
// ------------------------------------------------------------------------------------------------
template <int NW>
POA_DN void build_rows(Shared &sh, int qlen, int banded) {
    constexpr int NT = NW * POA_WARP;
    Ws &w = sh.ws;
    const int tid = poa_tid();
    const int n = sh.n_node;
    // four rows per thread in flight: every stage is a batch of independent loads (the walk is latency bound)
    for (int i0 = tid; i0 < n; i0 += 4 * NT) {
        int v[4], in[4], ioff[4], on[4], ooff[4], bs[4], rm[4], fpid[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int i = i0 + u * NT; v[u] = i < n ? w.idx2id[i] : -1; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            in[u] = on[u] = ioff[u] = ooff[u] = bs[u] = rm[u] = 0;
            if (v[u] >= 0) {
                in[u] = w.in_n[v[u]]; ioff[u] = w.in_off[v[u]]; on[u] = w.out_n[v[u]]; ooff[u] = w.out_off[v[u]];
                bs[u] = w.base[v[u]]; if (banded) rm[u] = w.remain[v[u]];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) fpid[u] = (v[u] >= 0 && in[u] > 0) ? w.pool_id[ioff[u]] : -1;
#pragma unroll
        for (int u = 0; u < 4; ++u) if (fpid[u] >= 0) fpid[u] = w.id2idx[fpid[u]];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (v[u] < 0) continue;
            const int i = i0 + u * NT;
            w.rowinfo[i] = poa_make_int4(ioff[u], in[u], ooff[u], on[u]);
            w.rbase[i] = (uint8_t)bs[u];
            w.tmp0[i] = fpid[u] >= 0 ? fpid[u] : 0;  // backtrack(), fill_p16(): row of the first predecessor
            if (in[u] > 0) w.pool_row[ioff[u]] = fpid[u];
            int4 p4 = poa_make_int4(in[u] > 0 ? fpid[u] : -1, -1, -1, -1);
            for (int k = 1; k < in[u]; ++k) {
                const int pr = w.id2idx[w.pool_id[ioff[u] + k]];
                w.pool_row[ioff[u] + k] = pr;
                if (k == 1) p4.y = pr; else if (k == 2) p4.z = pr; else if (k == 3) p4.w = pr;
            }
            w.pred4[i] = p4;  // multi-warp packed fill: the first four predecessors' rows, prefetched with the row's metadata
            w.tmp1[i] = p4.y;  // fill_p16(): row of the second predecessor, -1 if none
            for (int k = 0; k < on[u]; ++k) w.pool_row[ooff[u] + k] = w.id2idx[w.pool_id[ooff[u] + k]];
            if (banded) {
                w.rr[i] = qlen - rm[u];       // qlen - (remain[v] - remain[sink] - 1), remain[sink] = -1
                w.mplr[i] = n; w.mprr[i] = 0;  // abpoa_graph.c:349-352
            }
        }
    }
    sync_block<NW>();
}

// ------------------------------------------------------------------------------------------------
// DP fill
// ------------------------------------------------------------------------------------------------
template <typename S> struct VecIO;
template <> struct VecIO<short> {
    POA_SM void load(const short *p, int v[8]) {
        uint4 u = *reinterpret_cast<const uint4 *>(p);
        v[0] = (short)(u.x & 0xffffu); v[1] = (short)(u.x >> 16); v[2] = (short)(u.y & 0xffffu); v[3] = (short)(u.y >> 16);
        v[4] = (short)(u.z & 0xffffu); v[5] = (short)(u.z >> 16); v[6] = (short)(u.w & 0xffffu); v[7] = (short)(u.w >> 16);
    }
    POA_SM void store(short *p, const int v[8]) {
        uint4 u;
        u.x = ((unsigned)v[0] & 0xffffu) | ((unsigned)v[1] << 16); u.y = ((unsigned)v[2] & 0xffffu) | ((unsigned)v[3] << 16);
        u.z = ((unsigned)v[4] & 0xffffu) | ((unsigned)v[5] << 16); u.w = ((unsigned)v[6] & 0xffffu) | ((unsigned)v[7] << 16);
        *reinterpret_cast<uint4 *>(p) = u;
    }
};
template <> struct VecIO<int> {
    POA_SM void load(const int *p, int v[8]) {
        int4 a = *reinterpret_cast<const int4 *>(p), b = *reinterpret_cast<const int4 *>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    POA_SM void store(int *p, const int v[8]) {
        *reinterpret_cast<int4 *>(p) = poa_make_int4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<int4 *>(p + 4) = poa_make_int4(v[4], v[5], v[6], v[7]);
    }
};

template <typename S>
POA_D int inf_min_of(const DevParams &P) {  // abpoa_align_simd.c:1295 / :1299
    long long bm = sizeof(S) == 2 ? (long long)INT16_MIN : (long long)INT32_MIN;
    long long a = bm + P.min_mis, b = bm + P.oe1, c = bm + P.oe2;
    long long m = a > b ? a : b; m = m > c ? m : c;
    return (int)(m + 512 * (P.e1 > P.e2 ? P.e1 : P.e2));
}

// Row r of the current alignment is stored band-only: rowmeta[r] = {first vector index in the slab,
// beg, end, 0}; the stored planes (H, E1, E2 for convex gaps; H, E1 affine; H linear -- never the F planes) of
// nv = (end>>3)-(beg>>3)+1 vectors each follow one another.  Cells of those vectors that lie outside [beg,end] hold
// inf_min in every plane.
template <typename S>
POA_D const S *cell_ptr(const Ws &w, const int4 &pm, int plane, int j) {
    int vb = pm.y >> 3, nv = (pm.z >> 3) - vb + 1;
    return reinterpret_cast<const S *>(w.slab) + ((long long)pm.x + (long long)plane * nv) * 8 + (j - vb * 8);
}

// MODE: abpoa_set_gap_mode (abpoa_align.c:87-91) -- 0 convex (the reference's five planes H,E1,E2,F1,F2; simd_abpoa_cg_dp
// abpoa_align_simd.c:935-1074), 1 affine (H,E1,F1; simd_abpoa_ag_dp :817-933), 2 linear (H; simd_abpoa_lg_dp :727-815).
// Only H and the E planes are stored here (see NPL below).
// What the affine kernel does differently from "convex with one piece" is restated on purpose: F1 is fed by
// M + profile alone (:908), the stored E1 is reset where F1 strictly won the cell (:926,:930), local mode does
// not clamp E1, and the first cell of a row's first vector stores F1 = (M + profile) - oe1 of its own column
// (:898,:908).  The linear kernel folds the deletion into the predecessor pass and clamps local scores only
// after the horizontal pass (:813).
template <int NW, typename S, int MODE>
POA_DN void fill(Shared &sh, const DevParams &P, const uint8_t *q, int qlen, long long slab_vecs) {
    // planes STORED per row: H, E1, E2 (convex) / H, E1 (affine) / H (linear).  The F planes are evaluated but never stored:
    // no later row reads them, and the traceback recomputes them for one row on its insertion steps (generic_row_f()).
    constexpr int NPL = MODE == 0 ? 3 : (MODE == 1 ? 2 : 1);
    constexpr int NT = NW * POA_WARP;
    Ws &w = sh.ws;
    const int tid = poa_tid();
    const int lane = tid % POA_WARP, wid = tid / POA_WARP;
    const int n_node = sh.n_node;
    const int rows = n_node - 1;  // the sink row is never filled
    const int inf_min = inf_min_of<S>(P);
    const int pn = sizeof(S) == 2 ? P.pn16 : P.pn32;
    const int local = P.local;
    const int wb = local ? -1 : P.wb;  // abpoa_align.c:158
#ifdef POA_HOST_EMU
    const int bw = wb < 0 ? qlen : wb + (int)(P.wf * qlen);  // abpoa_align_simd.c:474
#else
    const int bw = wb < 0 ? qlen : wb + (int)__fmul_rn(P.wf, (float)qlen);
#endif
    const int e1 = P.e1, e2 = P.e2, oe1 = P.oe1, oe2 = P.oe2;
    S *slab = reinterpret_cast<S *>(w.slab);
    long long used = 0;  // vectors
    long long inband = 0, edge_rows = 0;
    int xbuf = 0;        // parity of the cross-warp exchange buffers
    (void)lane; (void)wid;

    // ---- row 0 (abpoa_align_simd.c:617-688)
    {
        int end0;
        if (wb >= 0) {
            if (tid == 0) {
                w.mplr[0] = 0; w.mprr[0] = 0;
                int4 ri = w.rowinfo[0];
                for (int k = 0; k < ri.w; ++k) { int o = w.pool_row[ri.z + k]; w.mplr[o] = 1; w.mprr[o] = 1; }
            }
            end0 = imin(qlen, imax(0, w.rr[0]) + bw);
        } else end0 = qlen;
        int nv = (end0 >> 3) + 1;
        if ((long long)NPL * nv > slab_vecs) { if (tid == 0) sh.err = ST_ESLAB; sync_block<NW>(); return; }
        if (tid == 0) w.rowmeta[0] = poa_make_int4(0, 0, end0, 0);
        for (int vc = tid; vc < nv; vc += NT) {
            int H[8], E1[8], E2[8], F1[8], F2[8];
            for (int c = 0; c < 8; ++c) {
                int j = vc * 8 + c;
                if (j > end0) { H[c] = E1[c] = E2[c] = F1[c] = F2[c] = inf_min; }
                else if (local) { H[c] = E1[c] = E2[c] = F1[c] = F2[c] = 0; }
                else if (j == 0) { H[c] = 0; E1[c] = (S)(-oe1); E2[c] = (S)(-oe2); F1[c] = F2[c] = inf_min; }
                else if (MODE == 0) {
                    F1[c] = (S)(-P.o1 - e1 * j); F2[c] = (S)(-P.o2 - e2 * j);
                    H[c] = imax(F1[c], F2[c]); E1[c] = E2[c] = inf_min;
                } else if (MODE == 1) { F1[c] = H[c] = (S)(-P.o1 - e1 * j); E1[c] = inf_min; E2[c] = F2[c] = inf_min; }
                else { H[c] = (S)(-e1 * j); E1[c] = E2[c] = F1[c] = F2[c] = inf_min; }
                if (MODE == 2 && j == 0 && !local) H[c] = 0;
            }
            S *p = slab + (long long)vc * 8;
            VecIO<S>::store(p, H);
            if (MODE == 0) {
                VecIO<S>::store(p + (long long)nv * 8, E1); VecIO<S>::store(p + 2LL * nv * 8, E2);
            } else if (MODE == 1) VecIO<S>::store(p + (long long)nv * 8, E1);
        }
        used = (long long)NPL * nv;
        inband += end0 + 1;
        sync_block<NW>();
    }
    int best_score = inf_min, best_i = 0, best_j = 0;
    const int f0_1 = imax((int)(S)(inf_min - oe1), (int)(S)(inf_min - e1));
    const int f0_2 = imax((int)(S)(inf_min - oe2), (int)(S)(inf_min - e2));

    // ---- rows in index order (abpoa_align_simd.c:1205-1221)
    for (int i = 1; i < rows; ++i) {
        const int4 ri = w.rowinfo[i];  // {in_off, in_n, out_off, out_n}
        const int *mrow = P.mat + 5 * w.rbase[i];
        int beg, end;
        if (wb < 0) { beg = 0; end = qlen; }
        else {  // abpoa_align.h:34-35, abpoa_align_simd.c:946-960
            int r = w.rr[i];
            beg = imax(0, imin(w.mplr[i], r) - bw);
            end = imin(qlen, imax(w.mprr[i], r) + bw);
            int beg_sn = beg / pn, min_pre_beg = INT_MAX, min_pre_beg_sn = INT_MAX;
            for (int k = 0; k < ri.y; ++k) {
                int pb = w.rowmeta[w.pool_row[ri.x + k]].y;
                if (min_pre_beg > pb) { min_pre_beg = pb; min_pre_beg_sn = pb / pn; }
            }
            if (beg_sn < min_pre_beg_sn) beg = min_pre_beg;
        }
        if (end < beg) end = beg;
        const int vb = beg >> 3, ve = end >> 3, nv = ve - vb + 1;
        if (used + (long long)NPL * nv > slab_vecs) { if (tid == 0) sh.err = ST_ESLAB; sync_block<NW>(); return; }
        const long long roff = used;
        used += (long long)NPL * nv;
        inband += end - beg + 1;
        edge_rows += (long long)ri.y * (end - beg + 1);

        // G[beg] of the gap pieces.  Affine: the scan starts from inf_min - oe1 (cell beg itself is patched below);
        // linear: nothing enters from the left of the band.
        int carry1 = MODE == 0 ? f0_1 + e1 * beg : (MODE == 1 ? (int)(S)(inf_min - oe1) + e1 * beg : NEG_INF32);
        int carry2 = f0_2 + e2 * beg;
        int mx = inf_min, left = -1, right = -1;
        for (int vc0 = vb; vc0 <= ve; vc0 += NT) {
            const int vc = vc0 + tid;
            const bool act = vc <= ve;
            const int j0 = vc * 8;
            int M[8], E1[8], E2[8];
            for (int c = 0; c < 8; ++c) { M[c] = inf_min; E1[c] = inf_min; E2[c] = inf_min; }
            if (act) {
                for (int k = 0; k < ri.y; ++k) {  // predecessors in in_id order (abpoa_align_simd.c:966-1029)
                    const int4 pm = w.rowmeta[w.pool_row[ri.x + k]];
                    const int pvb = pm.y >> 3, pve = pm.z >> 3, pnv = pve - pvb + 1;
                    const S *pH = slab + (long long)pm.x * 8;
                    if (vc >= pvb && vc <= pve) {
                        int hv[8], ev[8];
                        const S *ph = pH + (long long)(vc - pvb) * 8;
                        VecIO<S>::load(ph, hv);
                        for (int c = 0; c < 7; ++c) M[c + 1] = imax(M[c + 1], hv[c]);
                        if (MODE == 2) {  // linear: the deletion operand is the predecessor's H of the same column
                            for (int c = 0; c < 8; ++c) E1[c] = imax(E1[c], hv[c]);
                        } else {
                            VecIO<S>::load(ph + (long long)pnv * 8, ev);
                            for (int c = 0; c < 8; ++c) E1[c] = imax(E1[c], ev[c]);
                        }
                        if (MODE == 0) {
                            VecIO<S>::load(ph + 2LL * pnv * 8, ev);
                            for (int c = 0; c < 8; ++c) E2[c] = imax(E2[c], ev[c]);
                        }
                    }
                    if (vc - 1 >= pvb && vc - 1 <= pve) M[0] = imax(M[0], (int)pH[(long long)(vc - 1 - pvb) * 8 + 7]);
                }
                if (local && vc == 0) M[0] = imax(M[0], 0);  // abpoa_align_simd.c:974 (`first` = 0)
            }
            // H = M + profile; Hh = max(H, E1, E2); scan inputs (abpoa_align_simd.c:1032-1059)
            int Hh[8], c1[8], c2[8];
            int t1 = NEG_INF32, t2 = NEG_INF32;
            for (int c = 0; c < 8; ++c) {
                const int j = j0 + c;
                const bool inb = act && j >= beg && j <= end;
                int hh = inf_min;
                if (inb) {
                    int s = j == 0 ? 0 : mrow[q[j - 1]];
                    int hm = (S)(M[c] + s);
                    // exclusive prefixes: value seen by cell j is the max over cells < j
                    c1[c] = t1; c2[c] = t2;
                    if (MODE == 0) {
                        hh = imax(imax(hm, E1[c]), E2[c]);
                        t1 = imax(t1, hh - oe1 + e1 * (j + 1));
                        t2 = imax(t2, hh - oe2 + e2 * (j + 1));
                    } else if (MODE == 1) {
                        hh = hm;  // F1 is fed by M + profile alone; E1 joins after the scan
                        t1 = imax(t1, hm - oe1 + e1 * (j + 1));
                    } else {
                        hh = imax(hm, (int)(S)(E1[c] - e1));  // match/mismatch or deletion
                        t1 = imax(t1, hh - e1 + e1 * (j + 1));
                    }
                } else { c1[c] = t1; c2[c] = t2; }
                Hh[c] = hh;
            }
            // exclusive max-scan of the per-thread totals across the chunk
            int i1 = t1, i2 = t2;
            for (int d = 1; d < POA_WARP; d <<= 1) {
                int u1 = poa_shfl_up(i1, d), u2 = poa_shfl_up(i2, d);
                if (lane >= d) { i1 = imax(i1, u1); i2 = imax(i2, u2); }
            }
            int x1 = poa_shfl_up(i1, 1), x2 = poa_shfl_up(i2, 1);
            if (lane == 0) { x1 = NEG_INF32; x2 = NEG_INF32; }
            int tot1 = poa_shfl(i1, POA_WARP - 1), tot2 = poa_shfl(i2, POA_WARP - 1);
            if (NW > 1) {
                if (lane == POA_WARP - 1) { sh.scan_x[xbuf][0][wid] = i1; sh.scan_x[xbuf][1][wid] = i2; }
                poa_sync_block();
                int p1 = NEG_INF32, p2 = NEG_INF32, a1 = NEG_INF32, a2 = NEG_INF32;
                for (int ww = 0; ww < NW; ++ww) {
                    int v1 = sh.scan_x[xbuf][0][ww], v2 = sh.scan_x[xbuf][1][ww];
                    if (ww < wid) { p1 = imax(p1, v1); p2 = imax(p2, v2); }
                    a1 = imax(a1, v1); a2 = imax(a2, v2);
                }
                x1 = imax(x1, p1); x2 = imax(x2, p2); tot1 = a1; tot2 = a2;
                xbuf ^= 1;
            }
            x1 = imax(x1, carry1); x2 = imax(x2, carry2);
            carry1 = imax(carry1, tot1); carry2 = imax(carry2, tot2);
            if (act) {
                int H[8], F1[8], F2[8];
                for (int c = 0; c < 8; ++c) {
                    const int j = j0 + c;
                    if (j >= beg && j <= end) {
                        int f1, f2 = inf_min, h, ne1, ne2 = inf_min;
                        if (MODE == 0) {
                            f1 = (S)(imax(x1, c1[c]) - e1 * j);
                            f2 = (S)(imax(x2, c2[c]) - e2 * j);
                            h = imax(Hh[c], imax(f1, f2));
                            if (local) h = imax(h, 0);
                            ne1 = imax((int)(S)(E1[c] - e1), (int)(S)(h - oe1));
                            ne2 = imax((int)(S)(E2[c] - e2), (int)(S)(h - oe2));
                            if (local) { ne1 = imax(ne1, 0); ne2 = imax(ne2, 0); }
                        } else if (MODE == 1) {
                            f1 = (S)(imax(x1, c1[c]) - e1 * j);
                            if (j == beg) f1 = (beg % pn == 0) ? (int)(S)(Hh[c] - oe1) : (int)(S)(inf_min - oe1);  // abpoa_align_simd.c:898,:908
                            const int t = imax(Hh[c], E1[c]);
                            h = imax(t, f1);
                            if (local) h = imax(h, 0);
                            ne1 = h == t ? imax((int)(S)(E1[c] - e1), (int)(S)(h - oe1)) : (local ? 0 : inf_min);  // :926,:930
                        } else {
                            const int fl = imax(x1, c1[c]);
                            f1 = fl == NEG_INF32 ? inf_min : (int)(S)(fl - e1 * j);
                            h = imax(Hh[c], f1);
                            if (local) h = imax(h, 0);
                            ne1 = inf_min;
                        }
                        H[c] = h; F1[c] = f1; F2[c] = f2; E1[c] = ne1; E2[c] = ne2;
                        if (h > mx) { mx = h; left = right = j; } else if (h == mx) right = j;  // abpoa_align_simd.c:1107-1119
                    } else { H[c] = E1[c] = E2[c] = F1[c] = F2[c] = inf_min; }
                }
                S *p = slab + (roff + (vc - vb)) * 8;
                VecIO<S>::store(p, H);
                if (MODE == 0) {
                    VecIO<S>::store(p + (long long)nv * 8, E1); VecIO<S>::store(p + 2LL * nv * 8, E2);
                } else if (MODE == 1) VecIO<S>::store(p + (long long)nv * 8, E1);
                (void)F1; (void)F2;
            }
        }
        if (tid == 0) w.rowmeta[i] = poa_make_int4((int)roff, beg, end, 0);
        if (local || wb >= 0) {
            int amx = poa_redux_max(mx);
            int al = poa_redux_min((mx == amx && left >= 0) ? left : INT_MAX);
            int ar = poa_redux_max(mx == amx ? right : -1);
            if (NW > 1) {
                if (lane == 0) { sh.red_x[xbuf][0][wid] = amx; sh.red_x[xbuf][1][wid] = al; sh.red_x[xbuf][2][wid] = ar; }
                poa_sync_block();
                int gm = sh.red_x[xbuf][0][0];
                for (int ww = 1; ww < NW; ++ww) gm = imax(gm, sh.red_x[xbuf][0][ww]);
                int gl = INT_MAX, gr = -1;
                for (int ww = 0; ww < NW; ++ww) if (sh.red_x[xbuf][0][ww] == gm) {
                    gl = imin(gl, sh.red_x[xbuf][1][ww]); gr = imax(gr, sh.red_x[xbuf][2][ww]);
                }
                amx = gm; al = gl; ar = gr;
                xbuf ^= 1;
            }
            if (al == INT_MAX) al = -1;
            if (local && amx > best_score) { best_score = amx; best_i = i; best_j = al; }  // abpoa_align_simd.c:1208-1210
            if (wb >= 0) {  // abpoa_align_simd.c:1121-1130
                for (int k = tid; k < ri.w; k += NT) {
                    int o = w.pool_row[ri.z + k];
                    if (ar + 1 > w.mprr[o]) w.mprr[o] = ar + 1;
                    if (al + 1 < w.mplr[o]) w.mplr[o] = al + 1;
                }
            }
        }
        sync_block<NW>();
    }
    // ---- global best (abpoa_align_simd.c:1092-1105)
    if (tid == 0) {
        if (!local) {
            const int4 ri = w.rowinfo[rows];  // sink row
            for (int k = 0; k < ri.y; ++k) {
                int pi = w.pool_row[ri.x + k];
                const int4 pm = w.rowmeta[pi];
                int e = qlen > pm.z ? pm.z : qlen;
                int sc = *cell_ptr<S>(w, pm, 0, e);
                if (sc > best_score) { best_score = sc; best_i = pi; best_j = e; }
            }
        }
        sh.best_score = best_score; sh.best_i = best_i; sh.best_j = best_j;
        sh.inband += inband; sh.edge_rows += edge_rows;
    }
    sync_block<NW>();
}

// F1 / F2 of row i at columns j and j - 1 for rows of the generic fill (the traceback's insertion test,
// abpoa_align_simd.c:420-445 convex, :283-297 affine).  fill<>() evaluates F[j] = (S)(max(carry, max over in-band cells
// k < j of (Hh[k] - oe + e*(k+1))) - e*j) with Hh = max(M + profile, E1in, E2in) (convex) or M + profile (affine) and
// carry = f0 + e*beg; this re-evaluates exactly that expression for one row, the cells k dealt to the lanes of the
// traceback warp (predecessor cells straight from the slab), and a max-reduction in place of the scan.  The affine
// kernel's first band cell stores (M + profile) - oe1 of its own column (:898,:908).  out = {F1[j], F2[j], F1[j-1], F2[j-1]}
// (F2 = inf_min for affine; the j - 1 entries are inf_min when j - 1 < beg, as the traceback substitutes).
template <typename S, int MODE>
POA_DN void generic_row_f(Shared &sh, const DevParams &P, const uint8_t *q, int i, int j, int *out) {
    Ws &w = sh.ws;
    const int lane = poa_tid() % POA_WARP;
    const int inf_min = inf_min_of<S>(P);
    const int local = P.local;
    const int e1 = P.e1, e2 = P.e2, oe1 = P.oe1, oe2 = P.oe2;
    const int pn = sizeof(S) == 2 ? P.pn16 : P.pn32;
    const int4 rm = w.rowmeta[i], ri = w.rowinfo[i];
    const int beg = rm.y;
    const int *mrow = P.mat + 5 * w.rbase[i];
    const S *slab = reinterpret_cast<const S *>(w.slab);
    auto hh_at = [&](int k) {  // Hh of column k: what feeds the horizontal gaps of the cells to its right
        int M = inf_min, E1 = inf_min, E2 = inf_min;
        for (int t = 0; t < ri.y; ++t) {  // predecessors (abpoa_align_simd.c:966-1029)
            const int4 pm = w.rowmeta[w.pool_row[ri.x + t]];
            const int pvb = pm.y >> 3, pve = pm.z >> 3, pnv = pve - pvb + 1;
            const S *pH = slab + (long long)pm.x * 8;
            const int v = k >> 3;
            if (v >= pvb && v <= pve) {
                if (MODE != 2) E1 = imax(E1, (int)pH[((long long)(v - pvb) + pnv) * 8 + (k & 7)]);
                if (MODE == 0) E2 = imax(E2, (int)pH[((long long)(v - pvb) + 2LL * pnv) * 8 + (k & 7)]);
            }
            if (k >= 1) {
                const int vm = (k - 1) >> 3;
                if (vm >= pvb && vm <= pve) M = imax(M, (int)pH[(long long)(vm - pvb) * 8 + ((k - 1) & 7)]);
            }
        }
        if (local && k == 0) M = imax(M, 0);  // abpoa_align_simd.c:974
        const int sc = k == 0 ? 0 : mrow[q[k - 1]];
        const int hm = (S)(M + sc);
        return MODE == 0 ? imax(imax(hm, E1), E2) : hm;
    };
    int g1a = NEG_INF32, g1b = NEG_INF32, g2a = NEG_INF32, g2b = NEG_INF32;  // over k < j (a) and k < j - 1 (b)
    for (int k = beg + lane; k < j; k += POA_WARP) {
        const int hh = hh_at(k);
        const int c1 = hh - oe1 + e1 * (k + 1), c2 = hh - oe2 + e2 * (k + 1);
        g1a = imax(g1a, c1); g2a = imax(g2a, c2);
        if (k < j - 1) { g1b = imax(g1b, c1); g2b = imax(g2b, c2); }
    }
    g1a = poa_redux_max(g1a); g1b = poa_redux_max(g1b); g2a = poa_redux_max(g2a); g2b = poa_redux_max(g2b);
    const int f0_1 = imax((int)(S)(inf_min - oe1), (int)(S)(inf_min - e1));
    const int f0_2 = imax((int)(S)(inf_min - oe2), (int)(S)(inf_min - e2));
    const int carry1 = MODE == 0 ? f0_1 + e1 * beg : (int)(S)(inf_min - oe1) + e1 * beg;
    const int carry2 = f0_2 + e2 * beg;
    int F1j = (S)(imax(carry1, g1a) - e1 * j), F1l = inf_min, F2j = inf_min, F2l = inf_min;
    if (j - 1 >= beg) F1l = (S)(imax(carry1, g1b) - e1 * (j - 1));
    if (MODE == 0) {
        F2j = (S)(imax(carry2, g2a) - e2 * j);
        if (j - 1 >= beg) F2l = (S)(imax(carry2, g2b) - e2 * (j - 1));
    } else {  // affine: the row's first band cell
        if (j == beg) F1j = (beg % pn == 0) ? (int)(S)(hh_at(beg) - oe1) : (int)(S)(inf_min - oe1);
        if (j - 1 == beg) F1l = (beg % pn == 0) ? (int)(S)(hh_at(beg) - oe1) : (int)(S)(inf_min - oe1);
    }
    out[0] = F1j; out[1] = F2j; out[2] = F1l; out[3] = F2l;
}

#include "poa_fill16.cuh"

// ------------------------------------------------------------------------------------------------
// backtrack (abpoa_align_simd.c:309-458), one thread
// ------------------------------------------------------------------------------------------------
POA_D void push_cigar(Shared &sh, int op, int len, int node_id, int query_id) {  // abpoa_align.h:54-73
    unsigned long long *cig = sh.ws.cig + sh.cig_base;
    unsigned long long l = (unsigned long long)len;
    int n = sh.n_cigar;
    if (n == 0 || op != CINS || op != (int)(cig[n - 1] & 0xf)) {
        unsigned long long n_id = (unsigned long long)(long long)node_id, q_id = (unsigned long long)(long long)query_id;
        if (op == CMATCH) cig[n] = n_id << 34 | q_id << 4 | (unsigned)op;
        else if (op == CINS) cig[n] = q_id << 34 | l << 4 | (unsigned)op;
        else cig[n] = n_id << 34 | l << 4 | (unsigned)op;
        sh.n_cigar = n + 1;
    } else cig[n - 1] += l << 4;
}

// LAY16: rows are in the chunked layout of poa_fill16.cuh (S = short only)
template <typename S, bool LAY16>
POA_D const S *bt_cell(const Ws &w, const int4 &pm, int plane, int j) {
#if POA_WARP == 32
    if (LAY16) return reinterpret_cast<const S *>(cell_ptr16(w, pm, plane, j));
#endif
    return cell_ptr<S>(w, pm, plane, j);
}

// one iteration of the reference's traceback loop at cell (i, j) in state cur_op; 0 = moved, 1 = local alignment
// ends here (H == 0), 2 = dead end
// MODE 0 convex (abpoa_align_simd.c:309-458), 1 affine (:196-307: no second gap piece, F1 is plane 2), 2 linear (:116-194:
// match, then deletion, then insertion, no state)
// Rows hold no F planes: when the walk reaches the insertion test, bt_step() returns 3 without having changed anything,
// the warp recomputes {F1[j], F2[j], F1[j-1], F2[j-1]} of the row (p16_row_f() in poa_fill16.cuh for packed rows,
// generic_row_f() otherwise) and calls again with them in `fv`.
template <typename S, bool LAY16, int MODE>
POA_D int bt_step(Shared &sh, const DevParams &P, const uint8_t *q, int &i, int &j, int &id, int &cur_op, const int *fv = nullptr) {
    Ws &w = sh.ws;
    const int inf_min = inf_min_of<S>(P);
    const int local = P.local;
    const int e1 = P.e1, e2 = P.e2, oe1 = P.oe1, oe2 = P.oe2;
        const int4 rm = w.rowmeta[i];
        const int Hj = *bt_cell<S, LAY16>(w, rm, 0, j);
        if (local && Hj == 0) return 1;
        const int4 ri = w.rowinfo[i];
        const int s = P.mat[5 * w.rbase[i] + q[j - 1]];
        int hit = 0;
        if (MODE == 2 || (cur_op & OP_M)) {
            for (int k = 0; k < ri.y; ++k) {
                int pi = w.pool_row[ri.x + k];
                const int4 pm = w.rowmeta[pi];
                if (j - 1 < pm.y || j - 1 > pm.z) continue;
                if ((int)(S)(*bt_cell<S, LAY16>(w, pm, 0, j - 1) + s) == Hj) {
                    push_cigar(sh, CMATCH, 1, id, j - 1);
                    i = pi; --j; id = w.idx2id[i]; hit = 1; cur_op = OP_ALL;
                    break;
                }
            }
        }
        if (MODE == 2) {
            if (hit == 0) {
                for (int k = 0; k < ri.y; ++k) {
                    int pi = w.pool_row[ri.x + k];
                    const int4 pm = w.rowmeta[pi];
                    if (j < pm.y || j > pm.z) continue;
                    if ((int)(S)(*bt_cell<S, LAY16>(w, pm, 0, j) - e1) == Hj) {
                        push_cigar(sh, CDEL, 1, id, j - 1);
                        i = pi; id = w.idx2id[i]; hit = 1;
                        break;
                    }
                }
            }
            if (hit == 0) {
                const int hl = j - 1 >= rm.y ? (int)*bt_cell<S, LAY16>(w, rm, 0, j - 1) : inf_min;
                if ((int)(S)(hl - e1) == Hj) { push_cigar(sh, CINS, 1, id, j - 1); --j; hit = 1; }
            }
            if (hit == 0) { sh.err = ST_EINTERNAL; return 2; }
            return 0;
        }
        if (hit == 0 && (cur_op & OP_E)) {
            const int E1j = *bt_cell<S, LAY16>(w, rm, 1, j), E2j = MODE == 0 ? (int)*bt_cell<S, LAY16>(w, rm, 2, j) : inf_min;
            for (int k = 0; k < ri.y; ++k) {
                int pi = w.pool_row[ri.x + k];
                const int4 pm = w.rowmeta[pi];
                if (j < pm.y || j > pm.z) continue;
                const int pH = *bt_cell<S, LAY16>(w, pm, 0, j);
                if (cur_op & OP_E1) {
                    const int pE1 = *bt_cell<S, LAY16>(w, pm, 1, j);
                    bool cond = (cur_op & OP_M) ? (Hj == pE1) : (E1j == (int)(S)(pE1 - e1));
                    if (cond) {
                        cur_op = ((int)(S)(pH - oe1) == pE1) ? (OP_M | OP_F) : OP_E1;
                        hit = 1; push_cigar(sh, CDEL, 1, id, j - 1);
                        i = pi; id = w.idx2id[i];
                        break;
                    }
                }
                if (MODE == 0 && (cur_op & OP_E2)) {
                    const int pE2 = *bt_cell<S, LAY16>(w, pm, 2, j);
                    bool cond = (cur_op & OP_M) ? (Hj == pE2) : (E2j == (int)(S)(pE2 - e2));
                    if (cond) {
                        cur_op = ((int)(S)(pH - oe2) == pE2) ? (OP_M | OP_F) : OP_E2;
                        hit = 1; push_cigar(sh, CDEL, 1, id, j - 1);
                        i = pi; id = w.idx2id[i];
                        break;
                    }
                }
            }
        }
        if (hit == 0 && (cur_op & OP_F)) {
            if (fv == nullptr) return 3;  // no F planes are stored: the caller recomputes them for this row and calls again
            const bool inl = j - 1 >= rm.y;  // left neighbour inside this row's band?
            const int hl = inl ? (int)*bt_cell<S, LAY16>(w, rm, 0, j - 1) : inf_min;
            const int F1j = fv[0], F2j = MODE == 0 ? fv[1] : inf_min;
            if (cur_op & OP_F1) {
                if (!(cur_op & OP_M) || Hj == F1j) {
                    const int f1l = fv[2];
                    if ((int)(S)(hl - oe1) == F1j) { cur_op = OP_M | OP_E; hit = 1; }
                    else if ((int)(S)(f1l - e1) == F1j) { cur_op = OP_F1; hit = 1; }
                }
            }
            if (MODE == 0 && hit == 0 && (cur_op & OP_F2)) {
                if (!(cur_op & OP_M) || Hj == F2j) {
                    const int f2l = fv[3];
                    if ((int)(S)(hl - oe2) == F2j) { cur_op = OP_M | OP_E; hit = 1; }
                    else if ((int)(S)(f2l - e2) == F2j) { cur_op = OP_F2; hit = 1; }
                }
            }
            if (hit == 1) { push_cigar(sh, CINS, 1, id, j - 1); --j; }
        }
        if (hit == 0) { sh.err = ST_EINTERNAL; return 2; }
    return 0;
}

// Traceback by one warp.  The walk itself is sequential, but its dominant pattern is not: long diagonal runs
// of MATCH steps, each to the row's FIRST predecessor (the heaviest in-edge).  With cur_op = ALL the reference
// tries exactly that predecessor first (abpoa_align_simd.c:321-337), so lane t speculatively checks step t of
// such a run -- rows come from chasing the first-predecessor table fp[] -- and the leading run of successful
// lanes is committed at once; the first failing step falls back to bt_step() on lane 0.
template <typename S, bool LAY16, int MODE>
POA_DN void backtrack(Shared &sh, const DevParams &P, const uint8_t *q, int qlen) {
    Ws &w = sh.ws;
    const int lane = poa_tid() % POA_WARP;
    const int local = P.local;
    const int *fp = w.tmp0;  // build_rows(): row of the first predecessor
    int i = sh.best_i, j = sh.best_j;
    int id = w.idx2id[i];
    int cur_op = OP_ALL;
    if (lane == 0) { sh.n_cigar = 0; if (j < qlen) push_cigar(sh, CINS, qlen - j, -1, qlen - 1); }
    poa_sync_warp();
    int n = sh.n_cigar;
    poa_sync_warp();  // every lane has read it before lane 0 may write it again below
    bool skip_fast = false;
    while (i > 0 && j > 0) {
        if (POA_WARP > 1 && cur_op == OP_ALL && !skip_fast) {
            // rows of the run: chase the first-predecessor table (all lanes walk the same chain; 32 dependent but
            // cached loads -- rows of a bubble-rich graph are not consecutive, so the chain cannot be guessed).
            // MEASURED: serving the chain from a shared-memory window of fp[] changes nothing (backtrack stays 5.8 % of the
            // kernel): these loads hit L1, the phase's time is in the per-lane cell reads and the serial fallback steps.
            int ia = 0, ib = 0, r = i;
            for (int s = 0; s < POA_WARP; ++s) {
                const int nx = r > 0 ? fp[r] : 0;
                if (s == lane) { ia = r; ib = nx; }
                r = nx;
            }
            const int jt = j - lane;
            int ok = 0;
            if (ia > 0 && jt > 0) {                const int4 rm = w.rowmeta[ia], pm = w.rowmeta[ib];
                const int Hj = *bt_cell<S, LAY16>(w, rm, 0, jt);
                if (jt - 1 >= pm.y && jt - 1 <= pm.z && !(local && Hj == 0)) {
                    const int sc = P.mat[5 * w.rbase[ia] + q[jt - 1]];
                    ok = (int)(S)(*bt_cell<S, LAY16>(w, pm, 0, jt - 1) + sc) == Hj;
                }
            }
            const unsigned okm = poa_ballot(ok);
            const int nok = okm == 0xffffffffu ? 32 : p_ctz32(~okm);
            if (nok > 0) {
                if (lane < nok)  // abpoa_align.h:54-73: a MATCH always opens a new cigar word
                    (w.cig + sh.cig_base)[n + lane] = (unsigned long long)(long long)w.idx2id[ia] << 34 | (unsigned long long)(long long)(jt - 1) << 4 | (unsigned)CMATCH;
                n += nok;
                const int ni = poa_shfl(ib, nok - 1);  // first predecessor of the last committed row
                i = ni; j -= nok; id = w.idx2id[i]; cur_op = OP_ALL;
                skip_fast = nok < POA_WARP;  // the run ended on a cell that is not a first-predecessor match: the next attempt would commit nothing
                continue;
            }
        }
        if (lane == 0) {
            sh.n_cigar = n;
            int ti = i, tj = j, tid_ = id, top = cur_op;
            const int rc = bt_step<S, LAY16, MODE>(sh, P, q, ti, tj, tid_, top);
            sh.bcast[0] = ti; sh.bcast[1] = tj; sh.bcast[2] = top; sh.bcast[3] = rc;
        }
        poa_sync_warp();
        int rc = sh.bcast[3];
        if (rc == 3) {  // insertion test: the row's F planes are not stored; recompute them (all lanes), then redo the step
            poa_sync_warp();
            int fv[4];
#if POA_WARP == 32
            if (LAY16) p16_row_f(sh, P, qlen, i, j, fv); else
#endif
            generic_row_f<S, MODE>(sh, P, q, i, j, fv);
            if (lane == 0) {
                int ti = i, tj = j, tid_ = id, top = cur_op;
                const int rc2 = bt_step<S, LAY16, MODE>(sh, P, q, ti, tj, tid_, top, fv);
                sh.bcast[0] = ti; sh.bcast[1] = tj; sh.bcast[2] = top; sh.bcast[3] = rc2;
            }
            poa_sync_warp();
            rc = sh.bcast[3];
        }
        i = sh.bcast[0]; j = sh.bcast[1]; cur_op = sh.bcast[2]; n = sh.n_cigar;
        poa_sync_warp();
        if (rc == 2) return;
        if (rc == 1) break;
        id = w.idx2id[i];
        skip_fast = false;
    }
    if (lane == 0) {
        sh.n_cigar = n;
        if (j > 0) push_cigar(sh, CINS, j, -1, j - 1);
    }
    poa_sync_warp();
    n = sh.n_cigar;
    unsigned long long *cig = w.cig + sh.cig_base;
    for (int k = lane; k < n >> 1; k += POA_WARP) { unsigned long long t = cig[k]; cig[k] = cig[n - 1 - k]; cig[n - 1 - k] = t; }  // abpoa_align.h:88-96
    poa_sync_warp();
}

// ------------------------------------------------------------------------------------------------
// graph fusion (abpoa_graph.c:688-773), all threads.  path[qpos] = node id the base was placed on.  The reference
// walks the cigar once, serially: MATCH on the same base reuses the node, MATCH on another base takes the aligned
// node holding that base or creates one and links it into the aligned group, INS creates nodes, DEL adds nothing;
// every step adds `weight` to the edge from the previous path node (abpoa_add_graph_edge, :480-556), the last one
// to the sink.  Same result here without the serial walk: the read's path visits every graph node at most once,
// so each (node, edge list) pair is touched by exactly one query position and the positions can be
// processed independently once node ids are known; ids of created nodes are handed out in query order by
// a prefix sum, which is the order the reference creates them in.
// ------------------------------------------------------------------------------------------------
POA_D void edge_push_par(Shared &sh, int *off_arr, int *n_arr, int v, int id, int wt) {
    Ws &w = sh.ws;
    int n = n_arr[v], off = off_arr[v];
    if (n >= 2 && (n & (n - 1)) == 0) {
        int noff = poa_atomic_add(&sh.pool_used, 2 * n);
        if (noff + 2 * n > sh.pool_cap) { sh.err = ST_ESLAB; return; }
        for (int t = 0; t < n; ++t) { w.pool_id[noff + t] = w.pool_id[off + t]; w.pool_w[noff + t] = w.pool_w[off + t]; }
        off_arr[v] = noff; off = noff;
    }
    w.pool_id[off + n] = id; w.pool_w[off + n] = wt; n_arr[v] = n + 1;
}

template <int NW>
POA_D void block_incl_maxscan(Shared &sh, int *a, int n, int *scratch /* >= NT ints, global */) {
    constexpr int NT = NW * POA_WARP;
    const int tid = poa_tid();
    const int per = (n + NT - 1) / NT;
    const int lo = imin(n, tid * per), hi = imin(n, lo + per);
    int m = INT_MIN;
    for (int i = lo; i < hi; ++i) m = imax(m, a[i]);
    scratch[tid] = m;
    sync_block<NW>();
    int pre = INT_MIN;
    for (int t = 0; t < tid; ++t) pre = imax(pre, scratch[t]);
    sync_block<NW>();
    for (int i = lo; i < hi; ++i) { pre = imax(pre, a[i]); a[i] = pre; }
    sync_block<NW>();
}

template <int NW>
POA_DN int fuse_par(Shared &sh, const uint8_t *seq, int seq_l, int wt, int *path) {
    constexpr int NT = NW * POA_WARP;
    Ws &w = sh.ws;
    const int tid = poa_tid();
    const unsigned long long *cig = w.cig + sh.cig_base;
    const int n_cigar = sh.n_cigar;
    if (n_cigar == 0) return 0;  // abpoa_graph.c:706-708: the read is silently not added
    const int n_old = sh.n_node;
    int *flag = w.tmp0, *lastm = w.tmp1, *anc = w.tmp2, *cnode = w.tmp3, *scratch = w.rr;
    // cigar -> per query position: the graph node it is matched with, or -1 for an inserted base.  A MATCH op
    // carries its query index; an INS op carries the index of its last base and its length (abpoa_align.h:54-73).
    for (int i = tid; i < n_cigar; i += NT) {
        const unsigned long long c = cig[i];
        const int op = (int)(c & 0xf);
        if (op == CMATCH) cnode[(int)((c >> 4) & 0x3fffffff)] = (int)((c >> 34) & 0x3fffffff);
        else if (op == CINS) {
            const int qid = (int)((c >> 34) & 0x3fffffff), len = (int)((c >> 4) & 0x3fffffff);
            for (int j = 0; j < len; ++j) cnode[qid - j] = -1;
        }
    }
    sync_block<NW>();
    for (int t0 = tid; t0 < seq_l; t0 += 4 * NT) {  // four positions per thread in flight (latency-bound gathers)
        int x[4], b[4], xb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int t = t0 + u * NT; x[u] = t < seq_l ? cnode[t] : -2; b[u] = t < seq_l ? seq[t] : 0; }
#pragma unroll
        for (int u = 0; u < 4; ++u) xb[u] = x[u] >= 0 ? w.base[x[u]] : -1;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * NT;
            if (x[u] == -2) continue;
            int node = -1;
            if (x[u] >= 0) node = xb[u] == b[u] ? x[u] : get_aligned_id(w, x[u], b[u]);
            path[t] = node;              // existing node the base lands on, or -1: a node must be created
            flag[t] = node < 0;
            lastm[t] = x[u] >= 0 ? t : -1;
        }
    }
    sync_block<NW>();
    const int n_new = block_excl_scan<NW>(sh, flag, flag, seq_l, scratch);  // flag[t] = rank among created nodes
    if (n_old + n_new > sh.nmax) { sync_block<NW>(); if (tid == 0) sh.err = ST_ESLAB; sync_block<NW>(); return 0; }
    block_incl_maxscan<NW>(sh, lastm, seq_l, scratch);                      // last matched position <= t
    for (int t = tid; t < seq_l; t += NT) {
        if (path[t] >= 0) continue;
        const int id = n_old + flag[t], x = cnode[t];
        w.base[id] = seq[t]; w.aln_n[id] = 0;
        w.in_n[id] = 0; w.out_n[id] = 0; w.in_off[id] = 4 * id; w.out_off[id] = 4 * id + 2;
        if (x >= 0) { add_aligned(w, x, id); anc[flag[t]] = x; }
        else anc[flag[t]] = lastm[t] >= 0 ? cnode[lastm[t]] : SRC_ID;
        path[t] = id;
    }
    sync_block<NW>();
    if (tid == 0) sh.n_node = n_old + n_new;
    for (int t0 = tid; t0 <= seq_l; t0 += 4 * NT) {  // abpoa_graph.c:480-556; every (node, list) pair belongs to one position
        int from[4], to[4], inn[4], ino[4], outn[4], outo[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * NT;
            from[u] = t > seq_l ? -1 : (t == 0 ? SRC_ID : path[t - 1]);
            to[u] = t > seq_l ? -1 : (t == seq_l ? SINK_ID : path[t]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            inn[u] = ino[u] = outn[u] = outo[u] = 0;
            if (from[u] >= 0 && from[u] < n_old && to[u] < n_old) {
                inn[u] = w.in_n[to[u]]; ino[u] = w.in_off[to[u]]; outn[u] = w.out_n[from[u]]; outo[u] = w.out_off[from[u]];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (from[u] < 0) continue;
            int exist = 0;
            if (from[u] < n_old && to[u] < n_old) {
                for (int i = 0; i < inn[u]; ++i) if (w.pool_id[ino[u] + i] == from[u]) { w.pool_w[ino[u] + i] += wt; break; }
                for (int i = 0; i < outn[u]; ++i) if (w.pool_id[outo[u] + i] == to[u]) { w.pool_w[outo[u] + i] += wt; exist = 1; break; }
            }
            if (!exist) {
                edge_push_par(sh, w.in_off, w.in_n, to[u], from[u], wt);
                edge_push_par(sh, w.out_off, w.out_n, from[u], to[u], wt);
            }
        }
    }
    sync_block<NW>();
    return seq_l;
}

// ------------------------------------------------------------------------------------------------
// consensus (abpoa_output.c:468-536 + :375-391, one cluster) and MSA rank (abpoa_graph.c:359-410)
// ------------------------------------------------------------------------------------------------
POA_DN int heaviest_bundling(Shared &sh, int *cons) {  // one thread; uses tmp0..tmp3
    Ws &w = sh.ws;
    const int n = sh.n_node;
    int *outdeg = w.tmp0, *score = w.tmp1, *max_out = w.tmp2, *q = w.tmp3;
    for (int i = 0; i < n; ++i) { outdeg[i] = w.out_n[i]; max_out[i] = -1; score[i] = 0; }
    int qh = 0, qt = 0;
    q[qt++] = SINK_ID;
    while (qh < qt) {
        int cur = q[qh++];
        if (cur == SINK_ID) { max_out[cur] = -1; score[cur] = 0; }
        else {
            int max_id = -1;
            int on = w.out_n[cur], ooff = w.out_off[cur];
            if (cur == SRC_ID) {
                int path_score = -1, path_max_w = -1;
                for (int i = 0; i < on; ++i) {
                    int o = w.pool_id[ooff + i], ow = w.pool_w[ooff + i];
                    if (ow > path_max_w || (ow == path_max_w && score[o] > path_score)) { max_id = o; path_score = score[o]; path_max_w = ow; }
                }
                max_out[cur] = max_id;
                break;
            } else {
                int max_w = INT_MIN;
                for (int i = 0; i < on; ++i) {
                    int o = w.pool_id[ooff + i], ow = w.pool_w[ooff + i];
                    if (max_w < ow) { max_w = ow; max_id = o; }
                    else if (max_w == ow && score[max_id] <= score[o]) max_id = o;
                }
                score[cur] = max_w + score[max_id];
                max_out[cur] = max_id;
            }
        }
        int in = w.in_n[cur], ioff = w.in_off[cur];
        for (int i = 0; i < in; ++i) { int p = w.pool_id[ioff + i]; if (--outdeg[p] == 0) q[qt++] = p; }
    }
    int len = 0, cur = max_out[SRC_ID];
    while (cur != SINK_ID && cur >= 0) { cons[len++] = cur; cur = max_out[cur]; }
    return len;
}

POA_DN void set_msa_rank(Shared &sh, int *rank) {  // one thread; uses tmp0 (in-degree), tmp1 (stack)
    Ws &w = sh.ws;
    const int n = sh.n_node;
    int *indeg = w.tmp0, *st = w.tmp1;
    int sp = 0, msa_rank = 0;
    for (int i = 0; i < n; ++i) { indeg[i] = w.in_n[i]; rank[i] = -1; }
    st[sp++] = SRC_ID;
    while (sp > 0) {
        int cur = st[--sp];
        if (rank[cur] < 0) {
            rank[cur] = msa_rank;
            for (int i = 0; i < w.aln_n[cur]; ++i) rank[w.aln[4 * cur + i]] = msa_rank;
            msa_rank++;
        }
        if (cur == SINK_ID) break;
        int on = w.out_n[cur], ooff = w.out_off[cur];
        for (int i = 0; i < on; ++i) {
            int o = w.pool_id[ooff + i];
            if (--indeg[o] == 0) {
                int ok = 1, an = w.aln_n[o];
                for (int j = 0; j < an; ++j) if (indeg[w.aln[4 * o + j]] != 0) { ok = 0; break; }
                if (!ok) continue;
                st[sp++] = o; rank[o] = -1;
                for (int j = 0; j < an; ++j) { st[sp++] = w.aln[4 * o + j]; rank[w.aln[4 * o + j]] = -1; }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// one POA block, start to finish
// ------------------------------------------------------------------------------------------------
POA_D void ws_bind(Ws &w, char *b, const WsLayout &L) {
    w.base = (uint8_t *)(b + L.o_base); w.aln_n = (uint8_t *)(b + L.o_aln_n); w.aln = (int *)(b + L.o_aln);
    w.in_off = (int *)(b + L.o_in_off); w.in_n = (int *)(b + L.o_in_n); w.out_off = (int *)(b + L.o_out_off); w.out_n = (int *)(b + L.o_out_n);
    w.pool_id = (int *)(b + L.o_pool_id); w.pool_w = (int *)(b + L.o_pool_w); w.pool_row = (int *)(b + L.o_pool_row);
    w.idx2id = (int *)(b + L.o_idx2id); w.id2idx = (int *)(b + L.o_id2idx); w.remain = (int *)(b + L.o_remain);
    w.tmp0 = (int *)(b + L.o_tmp0); w.tmp1 = (int *)(b + L.o_tmp1); w.tmp2 = (int *)(b + L.o_tmp2); w.tmp3 = (int *)(b + L.o_tmp3);
    w.rowinfo = (int4 *)(b + L.o_rowinfo); w.rowmeta = (int4 *)(b + L.o_rowmeta); w.rbase = (uint8_t *)(b + L.o_rbase);
    w.rr = (int *)(b + L.o_rr); w.mplr = (int *)(b + L.o_mplr); w.mprr = (int *)(b + L.o_mprr); w.pred4 = (int4 *)(b + L.o_pred4);
    w.cig = (unsigned long long *)(b + L.o_cig); w.path = (int *)(b + L.o_path); w.best = (int *)(b + L.o_best); w.ncig = (int *)(b + L.o_ncig); w.plen = (int *)(b + L.o_plen); w.nrun = (int *)(b + L.o_nrun);
    w.qp = b + L.o_qp; w.slab = b + L.o_slab;
}

// block-wide sum of f(i), i in [0,n); result broadcast through sh.bcast[0..]
template <int NW, typename F>
POA_D long long block_sum(Shared &sh, int n, F f) {
    constexpr int NT = NW * POA_WARP;
    const int tid = poa_tid();
    int s = 0;
    for (int i = tid; i < n; i += NT) s += f(i);
    sync_block<NW>();
    if (tid == 0) sh.bcast[0] = 0;
    sync_block<NW>();
    if (s) poa_atomic_add(&sh.bcast[0], s);
    sync_block<NW>();
    int r = sh.bcast[0];
    sync_block<NW>();
    return r;
}

// exclusive prefix sum of src[0..n) into dst[0..n) (may alias), chunk per thread; returns total
template <int NW>
POA_D int block_excl_scan(Shared &sh, const int *src, int *dst, int n, int *scratch /* >= NT ints, global */) {
    constexpr int NT = NW * POA_WARP;
    const int tid = poa_tid();
    const int per = (n + NT - 1) / NT;
    const int lo = imin(n, tid * per), hi = imin(n, lo + per);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += src[i];
    scratch[tid] = s;
    sync_block<NW>();
    int pre = 0, tot = 0;
    for (int t = 0; t < NT; ++t) { int v = scratch[t]; if (t < tid) pre += v; tot += v; }
    sync_block<NW>();
    for (int i = lo; i < hi; ++i) { int v = src[i]; dst[i] = pre; pre += v; }
    sync_block<NW>();
    return tot;
}

template <int NW>
POA_D void poa_block(Shared &sh, const DevParams &P, const DevBatch &B, const WsLayout &L, const DevOut &O, int b, char *const wsb /* this CTA's workspace */) {
    constexpr int NT = NW * POA_WARP;
    Ws &w = sh.ws;
    const int tid = poa_tid();
    const long long s0 = B.block_seq_off[b], s1 = B.block_seq_off[b + 1];
    const int n_seq = (int)(s1 - s0);
    const long long base0 = B.seq_off[s0];
    const int banded = !P.local && P.wb >= 0;
    long long t_ph[PH_N];
    for (int k = 0; k < PH_N; ++k) t_ph[k] = 0;
    long long t_start = poa_clock();

    if (tid == 0) {
        sh.n_node = 0; sh.err = ST_OK; sh.inband = 0; sh.edge_rows = 0; sh.cig_base = 0; sh.n_cigar = 0;
        sh.nmax = L.nmax; sh.pool_cap = L.pool_cap;
        sh.pool_used = 4 * L.nmax;
        add_node(sh, 0); add_node(sh, 0);
        w.idx2id[0] = SRC_ID; w.idx2id[1] = SINK_ID; w.id2idx[SRC_ID] = 0; w.id2idx[SINK_ID] = 1;
    }
    sync_block<NW>();

    int block_lmax = 0;  // longest sequence of the block (p16_safe_for_long_graph)
    for (int k = 0; k < n_seq; ++k) block_lmax = imax(block_lmax, B.seq_len[s0 + k]);
    int cig_tot = 0;
    for (int k = 0; k < n_seq && sh.err == ST_OK; ++k) {
        const int qlen = B.seq_len[s0 + k];
        const long long qoff = B.seq_off[s0 + k] - base0;
        const uint8_t *q = B.bases + B.seq_off[s0 + k];
        const int wt = B.weight[s0 + k];
        int *path = w.path + qoff;
        int plen = 0;
        const int n_before = sh.n_node;
        if (tid == 0) { w.best[k] = 0; w.ncig[k] = 0; }
        if (sh.n_node == 2) {  // abpoa_align.c:193-198 / abpoa_graph.c:699-702
            long long t0 = poa_clock();
            add_first_sequence<NW>(sh, q, qlen, wt, path);
            plen = qlen;
            t_ph[PH_FUSE] += poa_clock() - t0;
        } else {
            long long t0 = poa_clock();
            build_rows<NW>(sh, qlen, banded);
            long long t1 = poa_clock();
            t_ph[PH_ROWS] += t1 - t0;
            // abpoa_align_simd.c:1286-1302: int16 unless the score range needs 32 bits
            const int gn = sh.n_node;
            const int len = qlen > gn ? qlen : gn;
            long long ms1 = (long long)qlen * P.match, ms2 = (long long)len * P.e1 + P.o1;
            long long max_score = ms1 > ms2 ? ms1 : ms2;
            const bool bits16 = max_score <= (long long)INT16_MAX - P.min_mis - P.oe1 - P.oe2;
#if POA_WARP == 32
            // packed 16-bit fill: whenever abPOA itself computes in int16, and beyond that while 16-bit cells provably hold every
            // real value (p16_safe_for_long_graph: deep blocks, whose graphs outgrow abPOA's int16 rule long before their scores do)
            const bool p16 = p16_eligible(P, qlen) && (bits16 || p16_safe_for_long_graph(P, qlen, block_lmax));
            const int p16_pn = bits16 ? P.pn16 : P.pn32;  // vector width of the reference kernel whose band-start rounding is reproduced
#else
            const bool p16 = false;
#endif
            if (p16) {
                t_ph[PH_SPARE] += 1;  // alignments that took the packed 16-bit fill
#if POA_WARP == 32
                if (NW == 1) {
                    if (P.p16_default) { if (P.local) fill_p16<NW, true, true>(sh, P, L, wsb, q, qlen, p16_pn); else fill_p16<NW, false, true>(sh, P, L, wsb, q, qlen, p16_pn); }
                    else { if (P.local) fill_p16<NW, true, false>(sh, P, L, wsb, q, qlen, p16_pn); else fill_p16<NW, false, false>(sh, P, L, wsb, q, qlen, p16_pn); }
                }
                else { if (P.local) fill_p16_mw<NW, true>(sh, P, q, qlen, L.slab_bytes, p16_pn); else fill_p16_mw<NW, false>(sh, P, q, qlen, L.slab_bytes, p16_pn); }
#endif
            } else if (P.gap_mode == 0) {
                if (bits16) fill<NW, short, 0>(sh, P, q, qlen, L.slab_bytes / 16); else fill<NW, int, 0>(sh, P, q, qlen, L.slab_bytes / 32);
            } else if (P.gap_mode == 1) {
                if (bits16) fill<NW, short, 1>(sh, P, q, qlen, L.slab_bytes / 16); else fill<NW, int, 1>(sh, P, q, qlen, L.slab_bytes / 32);
            } else {
                if (bits16) fill<NW, short, 2>(sh, P, q, qlen, L.slab_bytes / 16); else fill<NW, int, 2>(sh, P, q, qlen, L.slab_bytes / 32);
            }
            long long t2 = poa_clock();
            t_ph[PH_FILL] += t2 - t1;
            if (sh.err != ST_OK) break;
            if (tid < POA_WARP) {
                if (p16) backtrack<short, true, 0>(sh, P, q, qlen);
                else if (P.gap_mode == 0) { if (bits16) backtrack<short, false, 0>(sh, P, q, qlen); else backtrack<int, false, 0>(sh, P, q, qlen); }
                else if (P.gap_mode == 1) { if (bits16) backtrack<short, false, 1>(sh, P, q, qlen); else backtrack<int, false, 1>(sh, P, q, qlen); }
                else { if (bits16) backtrack<short, false, 2>(sh, P, q, qlen); else backtrack<int, false, 2>(sh, P, q, qlen); }
            }
            sync_block<NW>();
            long long t3 = poa_clock();
            t_ph[PH_BT] += t3 - t2;
            if (sh.err != ST_OK) break;
            if (tid == 0) { w.best[k] = sh.best_score; w.ncig[k] = sh.n_cigar; }
            plen = fuse_par<NW>(sh, q, qlen, wt, path);
            if (P.emit_cigar) cig_tot += sh.n_cigar;
            t_ph[PH_FUSE] += poa_clock() - t3;
            sync_block<NW>();
            if (tid == 0 && P.emit_cigar) sh.cig_base += sh.n_cigar;
        }
        if (tid == 0) w.plen[k] = plen;
        if (plen > 0 || sh.n_node == 2) {  // the reference re-sorts only when the read was added
            long long t0 = poa_clock();
            if (P.local) toposort<NW>(sh, banded); else toposort_incr<NW>(sh, banded, n_before);
            t_ph[PH_TOPO] += poa_clock() - t0;
        }
        sync_block<NW>();
    }
    sync_block<NW>();

    // ---- consensus, MSA rank, output
    long long tf0 = poa_clock();
    int *hdr = O.hdr + (long long)b * HDR_WORDS;
    const int n = sh.n_node;
    int status = sh.err;
    if (status == ST_OK) {
        int *plen_arr = w.plen;
        int cons_len = -1, msa_len = -1, msa_rows = 0;
        int *cons = w.mprr;
        int *rank = w.rr;
        if (tid == 0) {
            if (P.out_cons) cons_len = n > 2 ? heaviest_bundling(sh, cons) : 0;
            if (P.out_msa) {
                if (n > 2) {
                    set_msa_rank(sh, rank);
                    msa_len = rank[SINK_ID] - 1;
                    msa_rows = n_seq + (P.out_cons ? 1 : 0);
                    if (msa_len <= 0) { msa_len = 0; msa_rows = 0; }
                } else { msa_len = 0; msa_rows = 0; }
            }
            sh.bcast[1] = cons_len; sh.bcast[2] = msa_len; sh.bcast[3] = msa_rows;
        }
        sync_block<NW>();
        cons_len = sh.bcast[1]; msa_len = sh.bcast[2]; msa_rows = sh.bcast[3];
        sync_block<NW>();
        // section offsets
        int *in_pre = w.tmp0, *out_pre = w.tmp1, *aln_pre = w.tmp2, *scr = w.tmp3;
        const int in_tot = block_excl_scan<NW>(sh, w.in_n, in_pre, n, scr);
        const int out_tot = block_excl_scan<NW>(sh, w.out_n, out_pre, n, scr);
        for (int v = tid; v < n; v += NT) aln_pre[v] = w.aln_n[v];
        sync_block<NW>();
        const int aln_tot = block_excl_scan<NW>(sh, aln_pre, aln_pre, n, scr);
        long long path_tot = 0, wsum = 0;
        int maxlen = 0;
        for (int k = 0; k < n_seq; ++k) { path_tot += plen_arr[k]; wsum += B.weight[s0 + k]; maxlen = imax(maxlen, B.seq_len[s0 + k]); }
        const int format = (n < 65536 && wsum < 65536 && maxlen < 65536) ? WIRE_NARROW : WIRE_WIDE;
        const bool wide = format == WIRE_WIDE;
        // runs of consecutive node ids per read path (and of the consensus path): a warp walks a path 32 steps at a time
        int *nrun = w.nrun, *run_off = w.nrun + (n_seq + 2);
        const int wid_ = tid / POA_WARP, lane_ = tid % POA_WARP;
        auto path_of = [&](int k, int &len) -> const int * {
            if (k < n_seq) { len = plen_arr[k]; return w.path + (B.seq_off[s0 + k] - base0); }
            len = cons_len > 0 ? cons_len : 0; return cons;
        };
        for (int k = wid_; k <= n_seq; k += NW) {
            int len; const int *src = path_of(k, len);
            int cnt = 0, carry = 0;
            for (int t0 = 0; t0 < len; t0 += POA_WARP) {
                const int t = t0 + lane_;
                const int v = t < len ? src[t] : 0;
                int pv = poa_shfl_up(v, 1);
                if (lane_ == 0) pv = carry;
                cnt += poa_popc(poa_ballot(t < len && (t == 0 || v != pv + 1)));
                carry = poa_shfl(v, POA_WARP - 1);
            }
            if (lane_ == 0) nrun[k] = cnt;
        }
        sync_block<NW>();
        if (tid == 0) { int a = 0; for (int k = 0; k <= n_seq; ++k) { run_off[k] = a; a += nrun[k]; } run_off[n_seq + 1] = a; }
        sync_block<NW>();
        const long long run_tot = run_off[n_seq + 1];
        const long long msa_bytes = (long long)msa_rows * (msa_len > 0 ? msa_len : 0);
        WireLayout WL;
        wire_layout(WL, n, n_seq, in_tot, out_tot, aln_tot, run_tot, cig_tot, msa_bytes, format);
        const long long words = WL.words;
        if (tid == 0) {
            unsigned long long off = poa_atomic_add(O.arena_used, (unsigned long long)words);
            if (off + (unsigned long long)words > O.arena_cap) {
                sh.err = ST_EARENA;  // the host re-runs the block with an arena sized from this exact figure
                hdr[H_OFF_LO] = (int)(unsigned)((unsigned long long)words & 0xffffffffull); hdr[H_OFF_HI] = (int)(unsigned)((unsigned long long)words >> 32);
            }
            sh.bcast[0] = (int)(off & 0xffffffffull); sh.bcast[1] = (int)(off >> 32);
        }
        sync_block<NW>();
        status = sh.err;
        if (status == ST_OK) {
            const unsigned long long off = (unsigned long long)(unsigned)sh.bcast[0] | ((unsigned long long)(unsigned)sh.bcast[1] << 32);
            int *o = O.arena + off;
            uint8_t *o_base = (uint8_t *)(o + WL.o_base), *o_aln_n = (uint8_t *)(o + WL.o_aln_n);
            int *o_in_n = o + WL.o_in_n, *o_out_n = o + WL.o_out_n, *o_in_id = o + WL.o_in_id, *o_in_w = o + WL.o_in_w;
            int *o_out_id = o + WL.o_out_id, *o_out_w = o + WL.o_out_w, *o_aln_id = o + WL.o_aln_id;
            int *o_plen = o + WL.o_plen, *o_best = o + WL.o_best, *o_ncig = o + WL.o_ncig, *o_nrun = o + WL.o_nrun, *o_runs = o + WL.o_runs, *o_cig = o + WL.o_cig;
            uint8_t *o_msa = (uint8_t *)(o + WL.o_msa);
            // narrow or wide store of element idx of a section
            auto put = [&](int *sec, long long idx, int val) {
                if (wide) sec[idx] = val; else reinterpret_cast<unsigned short *>(sec)[idx] = (unsigned short)val;
            };
            // sections are padded to whole words: clear the pad bytes so that equal results are equal byte strings
            if (tid == 0) {
                if (n & 3) { o[WL.o_base + (n >> 2)] = 0; o[WL.o_aln_n + (n >> 2)] = 0; }
                if (!wide) {
                    if (n & 1) { o_in_n[n >> 1] = 0; o_out_n[n >> 1] = 0; }
                    if (in_tot & 1) { o_in_id[in_tot >> 1] = 0; o_in_w[in_tot >> 1] = 0; }
                    if (out_tot & 1) { o_out_id[out_tot >> 1] = 0; o_out_w[out_tot >> 1] = 0; }
                    if (aln_tot & 1) o_aln_id[aln_tot >> 1] = 0;
                }
            }
            sync_block<NW>();
            for (int v = tid; v < n; v += NT) {
                o_base[v] = w.base[v];
                int cn = w.in_n[v], off2 = w.in_off[v], dst = in_pre[v];
                put(o_in_n, v, cn);
                for (int k = 0; k < cn; ++k) { put(o_in_id, dst + k, w.pool_id[off2 + k]); put(o_in_w, dst + k, w.pool_w[off2 + k]); }
                cn = w.out_n[v]; off2 = w.out_off[v]; dst = out_pre[v];
                put(o_out_n, v, cn);
                for (int k = 0; k < cn; ++k) { put(o_out_id, dst + k, w.pool_id[off2 + k]); put(o_out_w, dst + k, w.pool_w[off2 + k]); }
                cn = w.aln_n[v]; dst = aln_pre[v];
                o_aln_n[v] = (uint8_t)cn;
                for (int k = 0; k < cn; ++k) put(o_aln_id, dst + k, w.aln[4 * v + k]);
            }
            // per-read paths and the consensus path as runs (only reads that were added have a path)
            for (int k = wid_; k <= n_seq; k += NW) {
                int len; const int *src = path_of(k, len);
                int at = run_off[k], carry = 0;
                for (int t0 = 0; t0 < len; t0 += POA_WARP) {
                    const int t = t0 + lane_;
                    const int v = t < len ? src[t] : 0;
                    int pv = poa_shfl_up(v, 1);
                    if (lane_ == 0) pv = carry;
                    const bool head = t < len && (t == 0 || v != pv + 1);
                    const unsigned hm = poa_ballot(head);
                    if (head) {
                        const long long r = at + poa_popc(hm & ((1u << lane_) - 1u));
                        put(o_runs, 2 * r, v); put(o_runs, 2 * r + 1, t);
                    }
                    at += poa_popc(hm);
                    carry = poa_shfl(v, POA_WARP - 1);
                }
                if (lane_ == 0) o_nrun[k] = nrun[k];
            }
            for (int k = tid; k < n_seq; k += NT) { o_plen[k] = plen_arr[k]; o_best[k] = w.best[k]; o_ncig[k] = P.emit_cigar ? w.ncig[k] : 0; }
            for (int t = tid; t < cig_tot; t += NT) { unsigned long long c = w.cig[t]; o_cig[2 * t] = (int)(unsigned)(c & 0xffffffffull); o_cig[2 * t + 1] = (int)(unsigned)(c >> 32); }
            if (msa_bytes > 0) {  // abpoa_output.c:149-192
                for (long long t = tid; t < ((msa_bytes + 3) & ~3LL); t += NT) o_msa[t] = t < msa_bytes ? 5 : 0;
                sync_block<NW>();
                for (int k = 0; k < n_seq; ++k) {
                    const int pl = plen_arr[k];
                    const int *src = w.path + (B.seq_off[s0 + k] - base0);
                    for (int t = tid; t < pl; t += NT) { int nd = src[t]; o_msa[(long long)k * msa_len + rank[nd] - 1] = w.base[nd]; }
                }
                if (P.out_cons) for (int t = tid; t < cons_len; t += NT) { int nd = cons[t]; o_msa[(long long)n_seq * msa_len + rank[nd] - 1] = w.base[nd]; }
            }
            if (tid == 0) {
                hdr[H_N_NODE] = n; hdr[H_N_SEQ] = n_seq; hdr[H_CONS_LEN] = cons_len; hdr[H_MSA_LEN] = msa_len; hdr[H_MSA_ROWS] = msa_rows;
                hdr[H_IN_TOT] = in_tot; hdr[H_OUT_TOT] = out_tot; hdr[H_ALN_TOT] = aln_tot; hdr[H_PATH_TOT] = (int)path_tot; hdr[H_CIG_TOT] = cig_tot;
                hdr[H_OFF_LO] = sh.bcast[0]; hdr[H_OFF_HI] = sh.bcast[1];
                hdr[H_INBAND_LO] = (int)(unsigned)(sh.inband & 0xffffffffll); hdr[H_INBAND_HI] = (int)(unsigned)((unsigned long long)sh.inband >> 32);
                hdr[H_EDGE_LO] = (int)(unsigned)(sh.edge_rows & 0xffffffffll); hdr[H_EDGE_HI] = (int)(unsigned)((unsigned long long)sh.edge_rows >> 32);
                hdr[H_BODY_WORDS] = (int)words; hdr[H_FORMAT] = format; hdr[H_RUN_TOT] = (int)run_tot;
            }
        }
    }
    if (tid == 0) {
        hdr[H_STATUS] = status;
        if (status != ST_OK) { hdr[H_N_NODE] = 0; hdr[H_N_SEQ] = n_seq; }
        long long t_end = poa_clock();
        t_ph[PH_FINAL] = t_end - tf0; t_ph[PH_TOTAL] = t_end - t_start;
        for (int k = 0; k < PH_N; ++k) if (t_ph[k]) poa_atomic_add(O.phase + k, (unsigned long long)t_ph[k]);
    }
    sync_block<NW>();
}

}  // namespace poa
