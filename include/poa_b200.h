/* poa_b200.h -- C ABI of the B200-native partial-order-alignment engine.
 *
 * This is the drop-in boundary for smoothxg's per-block POA hot path.  The reference has no
 * plugin/FFI layer: the engine is chosen by an `if` inside the OpenMP block loop
 * (reference src/smooth.cpp:2075-2119) and the abPOA arithmetic is reached through
 * smooth_abpoa (src/smooth.cpp:133-627).  The cut that leaves everything else untouched is
 * *inside* smooth_abpoa: this ABI replaces exactly src/smooth.cpp:256-351
 *   abpoa_init_para / abpoa_post_set_para      (src/smooth.cpp:256-297)
 *   abpoa_init / abpoa_reset / base encoding    (src/smooth.cpp:300-324)
 *   abpoa_poa                                    (src/smooth.cpp:337, deps/abPOA/src/abpoa_align.c:304-344)
 *   abpoa_generate_rc_msa                        (src/smooth.cpp:342-344, deps/abPOA/src/abpoa_output.c:149-192)
 *   abpoa_generate_consensus                     (src/smooth.cpp:346-351, deps/abPOA/src/abpoa_output.c:1281-1312)
 * and hands back, in flat arrays, every field of abpoa_t that the host side of smoothxg reads
 * afterwards (MSA trimming src/smooth.cpp:362-516, build_odgi_abPOA src/smooth.cpp:2442-2574):
 * node bases, in/out edge lists in abPOA's final (weight-sorted) order with weights, per-read node
 * paths (what build_odgi_abPOA decodes from the read_ids bitsets, src/smooth.cpp:2488-2501),
 * consensus node ids, and the row-column MSA.  Node ids are abPOA's: 0 = source, 1 = sink, creation
 * order thereafter.  INTEGRATION.md shows the patch a smoothxg maintainer would apply.
 *
 * Plain C, plain pointers and sizes.  No C++/torch types cross this boundary.  All entry points
 * are thread-safe with respect to distinct engines/batches; one engine may be shared by several
 * host threads (calls serialise on an internal mutex).
 *
 * Everything behind this ABI runs on the GPU (hand-written sm_100a kernels); there is no CPU
 * fallback.  Blocks the device could not finish (workspace exhausted even after the engine's
 * automatic retry with larger workspaces) come back with a non-zero
 * per-block status and the call returns POA_B200_EBLOCK.
 */
#ifndef POA_B200_H
#define POA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POA_B200_ABI_VERSION 1
#define POA_B200_HDR_WORDS 20

/* return / per-block status codes */
enum {
    POA_B200_OK        = 0,
    POA_B200_ESLAB     = 1, /* per-block: DP workspace exhausted (retried automatically with a larger one) */
    POA_B200_EARENA    = 2, /* per-block: result arena exhausted (retried automatically with a larger one) */
    POA_B200_EINTERNAL = 3, /* per-block: traceback reached a dead end (the reference aborts here, abpoa_align_simd.c:448) */
    POA_B200_EUNSUP    = 4, /* unsupported parameters (gap_ext1 == gap_ext2 == 0, scores outside the 16-bit head room abPOA assumes) */
    POA_B200_EBLOCK    = 5, /* call-level: at least one block has a non-zero status */
    POA_B200_ECUDA     = 6, /* CUDA runtime error; see poa_b200_last_error() */
    POA_B200_EARG      = 7, /* bad argument */
    POA_B200_ENOMEM    = 8  /* host or device allocation failed */
};

/* Mirrors the abpoa_para_t fields smooth_abpoa sets (reference src/smooth.cpp:256-297).
 * Penalties are positive numbers, as smoothxg passes them (src/smooth.cpp:282-287).  The gap mode follows
 * abpoa_set_gap_mode (deps/abPOA/src/abpoa_align.c:87-91): gap_open1 == 0 linear, gap_open2 == 0 affine (what a
 * four-value -p gives, src/main.cpp:353-359), otherwise convex (smoothxg's default and all -a presets). */
typedef struct poa_b200_params {
    int32_t match, mismatch, gap_open1, gap_ext1, gap_open2, gap_ext2;
    int32_t align_mode; /* 0 = global, 1 = local (src/smooth.cpp:259-263); local forces wb = -1 (abpoa_align.c:158) */
    int32_t wb;         /* 311 = adaptive band, -1 = unbanded (src/smooth.cpp:266-270) */
    float   wf;         /* 0.03 (src/smooth.cpp:271) */
    int32_t out_cons;   /* src/smooth.cpp:275 */
    int32_t out_msa;    /* src/smooth.cpp:278 */
} poa_b200_params_t;

/* Engine tuning; zero-initialise for defaults. */
typedef struct poa_b200_engine_opts {
    int32_t warps_per_block;   /* CUDA warps cooperating on one POA block: 1, 2, 4 or 8; 0 = choose from the batch size */
    int32_t ctas_per_sm;       /* resident POA blocks per SM; 0 = occupancy-derived default */
    int32_t emit_cigar;        /* 1: also return per-sequence graph cigars (debug / parity instrumentation) */
    int32_t flags;             /* bit 0: disable the packed 16-bit fill (A/B testing; results are identical either way);
                                * bits 4-5: SIMD width of the abPOA build whose vector-granular band-start rounding is reproduced
                                * (deps/abPOA/src/abpoa_align_simd.c:949-960): 0 = AVX-512BW (32 int16 lanes; the default, and what the
                                * golden vectors were pinned with), 1 = AVX2 (16), 2 = SSE4.1 / NEON (8) */
    double  slab_rows_factor;  /* DP workspace rows per query base before a retry is needed; 0 = default (1.7) */
    int64_t device_mem_budget; /* bytes of HBM the engine may use for workspaces; 0 = 70 % of free memory */
} poa_b200_engine_opts_t;

typedef struct poa_b200_engine poa_b200_engine_t;
typedef struct poa_b200_batch  poa_b200_batch_t;
typedef struct poa_b200_result poa_b200_result_t;

/* One block of a finished batch.  All pointers point into memory owned by the result.  Results are held in a compact form
 * (narrow integers, run-length coded paths: about 60 KB for a 32 x 2 kb block instead of 386 KB as flat int32 arrays -- this is
 * what crosses PCIe and NVLink); poa_b200_result_block() expands a block into the flat arrays below the first time it is asked
 * for and keeps the expansion until poa_b200_result_release_block() or poa_b200_result_free(). */
typedef struct poa_b200_block_view {
    int32_t status;              /* POA_B200_OK or a per-block error */
    int32_t n_node;              /* abg->node_n, including source (0) and sink (1) */
    int32_t n_seq;
    int32_t cons_len;            /* -1 when no consensus was requested */
    int32_t msa_len;             /* -1 when no MSA was requested */
    int32_t msa_rows;            /* n_seq (+1 when a consensus row is present) */
    const int32_t *base;         /* [n_node] 0..3 = ACGT, 4 = N  (abg->node[i].base) */
    const int32_t *in_n;         /* [n_node] in_edge_n */
    const int32_t *in_id;        /* concatenated in_id lists, node order */
    const int32_t *in_w;         /* concatenated in_edge_weight lists */
    const int32_t *out_n;        /* [n_node] out_edge_n */
    const int32_t *out_id;       /* concatenated out_id lists */
    const int32_t *out_w;        /* concatenated out_edge_weight lists */
    const int32_t *aln_n;        /* [n_node] aligned_node_n */
    const int32_t *aln_id;       /* concatenated aligned_node_id lists */
    const int32_t *path_len;     /* [n_seq] nodes on each read's path (0 if the read was not added, abpoa_graph.c:706-708) */
    const int32_t *path_node;    /* concatenated per-read node ids, read order */
    const int32_t *cons_node;    /* [max(cons_len,0)] abc->cons_node_ids[0] */
    const uint8_t *msa;          /* [msa_rows * msa_len] abc->msa_base, gap = 5 */
    const int32_t *best_score;   /* [n_seq] per-sequence alignment score (0 for the first sequence) */
    const int32_t *n_cigar;      /* [n_seq] cigar words per sequence (all 0 unless emit_cigar) */
    const uint64_t *cigar;       /* concatenated abpoa_cigar_t words (deps/abPOA/include/abpoa.h:46-51) */
    int64_t in_total, out_total, aln_total, path_total, cigar_total;
    int64_t inband_cells;        /* DP cells evaluated inside the band for this block */
} poa_b200_block_view_t;

typedef struct poa_b200_stats {
    double  kernel_ms;       /* device time of the batch's POA kernel launches, re-runs of overflowed blocks included (CUDA events on the launch stream) */
    double  h2d_ms, d2h_ms;
    int64_t inband_cells;    /* summed over blocks */
    int64_t edge_row_cells;  /* sum over evaluated rows of (predecessor count x band width): p-bar = this / inband_cells */
    int64_t h2d_bytes, d2h_bytes;
    int32_t kernel_launches; /* POA kernel launches, retries included */
    int32_t retried_blocks;
    int32_t n_ctas;          /* persistent CTAs of the main launch */
    int32_t warps_per_block;
    int64_t workspace_bytes;
    int64_t phase_cycles[8]; /* summed SM cycles per phase: rows, fill, backtrack, fuse, toposort, finalize, total, spare */
} poa_b200_stats_t;

int  poa_b200_abi_version(void);
const char *poa_b200_strerror(int code);
const char *poa_b200_last_error(void); /* thread-local text of the last CUDA/argument error */

/* ASCII -> abPOA base codes, replacing the ab_char26_table loop at src/smooth.cpp:304-313 (table:
 * deps/abPOA/src/abpoa_seq.c:15-32): A/a 0, C/c 1, G/g 2, T/t/U/u 3, bytes 0..3 map to themselves, everything
 * else (N, IUPAC codes, '-') 4.  Pure host code; needs no GPU. */
void poa_b200_encode_bases(const char *ascii, int64_t n, uint8_t *codes);

int  poa_b200_engine_create(int device, const poa_b200_engine_opts_t *opts, poa_b200_engine_t **out);
void poa_b200_engine_destroy(poa_b200_engine_t *eng);
/* Give the engine's pooled device and pinned host buffers back to the driver (they are otherwise kept between batches so
 * that steady-state calls do no cudaMalloc / cudaMallocHost).  Call it before another user of the same GPU needs the memory. */
int  poa_b200_engine_trim(poa_b200_engine_t *eng);

/* One-shot, host buffers in / host result out (what the smoothxg loop body calls once per batch):
 *   block_seq_off[n_blocks+1] -> index into seq_len/weight;  seq_off[n_seqs+1] -> index into bases;
 *   bases = codes 0..4 (ab_char26_table encoding, deps/abPOA/src/abpoa_seq.c:15-32); a code above 4 -- which that
 *   table cannot produce -- is treated as 4 (N), clamped on the device right after the upload;
 *   weight[n_seqs] = dedup multiplicity of each sequence (src/smooth.cpp:332-336).
 * With POA_B200_TRACE set in the environment every call prints the host wall time of its stages (upload, launch + wait,
 * download, free) to stderr. */
int  poa_b200_run_batch(poa_b200_engine_t *eng, const poa_b200_params_t *params, int64_t n_blocks,
                        const int64_t *block_seq_off, const int32_t *seq_len, const int64_t *seq_off,
                        const uint8_t *bases, const int32_t *weight, poa_b200_result_t **result);

/* Per-block entry points with abpoa_poa's own argument shapes (deps/abPOA/src/abpoa_align.c:304): seqs[i] = codes of sequence
 * i, weights[i] = its dedup multiplicity.  They are what the UNMODIFIED OpenMP loop over blocks (src/smooth.cpp:1931) would call
 * from its worker threads, and they coalesce: blocks submitted by any number of host threads accumulate in one pending batch
 * that the engine's dispatcher thread launches when it holds 8 192 blocks, when a block with other parameters arrives, or as
 * soon as some thread waits for one of its blocks while the GPU is idle; submissions arriving while a batch runs form the next
 * one.  So N concurrent callers share launches instead of taking turns on the GPU.
 *   poa_b200_submit_block  copies the block and returns a ticket at once;
 *   poa_b200_wait_block    blocks until that block is done and returns a one-block result (block index 0; it keeps the shared
 *                          batch result alive until poa_b200_result_free).  A ticket can be collected once, from any thread.
 *   poa_b200_poa_block     = submit + wait.  With T synchronous callers a launch carries at most T blocks -- enough to keep
 *                          the reference's loop body unchanged, not enough to fill a B200: throughput needs either
 *                          poa_b200_run_batch or a window of outstanding tickets per thread. */
int  poa_b200_submit_block(poa_b200_engine_t *eng, const poa_b200_params_t *params, int32_t n_seq,
                           const uint8_t *const *seqs, const int32_t *seq_lens, const int32_t *weights, uint64_t *ticket);
int  poa_b200_wait_block(poa_b200_engine_t *eng, uint64_t ticket, poa_b200_result_t **result);
int  poa_b200_poa_block(poa_b200_engine_t *eng, const poa_b200_params_t *params, int32_t n_seq,
                        const uint8_t *const *seqs, const int32_t *seq_lens, const int32_t *weights,
                        poa_b200_result_t **result);

/* Staged API (inputs resident in HBM; used for kernel-only timing and by multi-GPU drivers). */
int  poa_b200_batch_upload(poa_b200_engine_t *eng, const poa_b200_params_t *params, int64_t n_blocks,
                           const int64_t *block_seq_off, const int32_t *seq_len, const int64_t *seq_off,
                           const uint8_t *bases, const int32_t *weight, poa_b200_batch_t **batch);
/* Enqueue the POA kernel on `stream` (a cudaStream_t, NULL = the engine's own stream); asynchronous. */
int  poa_b200_batch_launch(poa_b200_batch_t *batch, void *stream);
/* Wait, re-run blocks that exhausted a workspace, copy the results to the host. */
int  poa_b200_batch_download(poa_b200_batch_t *batch, void *stream, poa_b200_result_t **result);
/* Wait for the launch and re-run overflowed blocks, results stay on the device (kernel-only timing). */
int  poa_b200_batch_finish(poa_b200_batch_t *batch, void *stream);
/* After poa_b200_batch_finish: the device-resident result, for a multi-GPU gather over NVLink without a
 * host round trip.  *d_hdr = [n_blocks][POA_B200_HDR_WORDS] int32 block headers; arena `arena_idx`
 * (0 <= arena_idx < *n_arenas; more than one only if blocks were re-run) holds *arena_words int32 words of
 * block bodies; block_arena[b] (host array, n_blocks entries, owned by the batch) says which arena holds
 * block b.  poa_b200_result_from_parts() turns gathered copies back into a result. */
int  poa_b200_batch_device_result(poa_b200_batch_t *batch, int32_t arena_idx, const int32_t **d_hdr, const int32_t **d_arena,
                                  int64_t *arena_words, int32_t *n_arenas, const int32_t **block_arena);
/* Build a host result from header and arena words (copied) as produced on any GPU. */
int  poa_b200_result_from_parts(int64_t n_blocks, const int32_t *hdr, const int32_t *arena, int64_t arena_words,
                                poa_b200_result_t **result);
/* Same, but the arena words are still in device memory on `eng`'s GPU (rank 0 after the NCCL gather of every rank's
 * bodies): ONE device-to-host copy on `stream` (NULL = the engine's) straight into the result's pooled pinned buffer, no
 * second host copy.  hdr is a host array with body offsets already rebased onto d_arena. */
int  poa_b200_result_from_device_parts(poa_b200_engine_t *eng, int64_t n_blocks, const int32_t *hdr, const int32_t *d_arena,
                                       int64_t arena_words, void *stream, poa_b200_result_t **result);
void poa_b200_batch_free(poa_b200_batch_t *batch);
int  poa_b200_batch_stats(const poa_b200_batch_t *batch, poa_b200_stats_t *stats);

/* ---- next stage: the per-block graph smoothxg builds from the POA result.
 * build_odgi_abPOA (reference src/smooth.cpp:2442-2574) turns abpoa_t into an odgi graph of 1-bp nodes
 * (odgi id = abPOA id - 1), embeds one path per read with `padding_len` steps trimmed at both ends, embeds the
 * consensus path restricted to nodes some read still covers, and drops every node no path uses together with its edges
 * (:2567-2573; the edge-dropping pass before it, :2559-2565, removes nothing: odgi's find_edges_exceeding_depth_limits with
 * min_depth 1 only inspects edges that a path walks, deps/odgi/src/algorithms/depth.cpp:17-51 -- so an edge between two kept
 * nodes stays even when no trimmed path walks it).  poa_b200_block_graph() produces exactly that graph as flat arrays (pure host code, no GPU):
 * nodes in the order build_odgi_abPOA creates them (Kahn walk from the source over out_id order, :2463-2511),
 * edges as (from, to) pairs of odgi ids in creation order, paths as step lists in the POA's forward
 * orientation -- for a read whose dup_is_revs flag is set the caller walks its list backwards and flips the
 * handles, as :2524-2529 does; duplicate names of a deduplicated read share its list. */
typedef struct poa_b200_graph_view {
    int32_t n_node;             /* nodes kept: covered by at least one trimmed read path */
    const int32_t *node_id;     /* [n_node] odgi id (abPOA id - 1) */
    const char    *node_base;   /* [n_node] 'A','C','G','T','N' (ab_nt256_table) */
    int32_t n_edge;             /* edges kept: both end nodes kept */
    const int32_t *edge_from;   /* [n_edge] odgi ids, forward strand both ends */
    const int32_t *edge_to;
    int32_t n_path;             /* n_seq, plus one when the consensus path was requested */
    const int64_t *path_off;    /* [n_path + 1] offsets into path_node; the consensus path is the last one */
    const int32_t *path_node;   /* odgi ids */
} poa_b200_graph_view_t;
typedef struct poa_b200_graph poa_b200_graph_t;
int  poa_b200_block_graph(const poa_b200_block_view_t *view, int32_t padding_len, int32_t include_consensus,
                          poa_b200_graph_t **graph);
int  poa_b200_graph_view(const poa_b200_graph_t *graph, poa_b200_graph_view_t *view);
void poa_b200_graph_free(poa_b200_graph_t *graph);

/* ---- and the graph smooth_abpoa finally returns (reference src/smooth.cpp:545-620): the block graph above after
 * odgi::algorithms::unchop (runs of 1-bp nodes that no path enters, leaves or branches inside become one node,
 * deps/odgi/src/algorithms/unchop.cpp, simple_components.cpp, perfect_neighbors.cpp), renumbered 1..n in a topological order
 * (apply_ordering(topological_order(..), compact), :557), with one edge per consecutive pair of path steps (:590-606) and the
 * paths re-expressed over the merged nodes.  Node ids are ranks of OUR deterministic topological order (Kahn, ready nodes in
 * creation order); the reference's ids come out of hash-map iteration order after unchop and are not a function of the block,
 * so the two graphs are equal up to that renumbering: same node sequences, same edges, same path walks.  Pure host code.
 * Paths as in poa_b200_block_graph: forward orientation, read order, the consensus path last. */
typedef struct poa_b200_final_graph_view {
    int32_t n_node;             /* node k (0-based) has id k + 1 */
    const int64_t *seq_off;     /* [n_node + 1] offsets into seq */
    const char    *seq;         /* concatenated node sequences */
    int32_t n_edge;
    const int32_t *edge_from;   /* [n_edge] node ids, forward strand both ends, sorted */
    const int32_t *edge_to;
    int32_t n_path;
    const int64_t *path_off;    /* [n_path + 1] */
    const int32_t *path_node;   /* node ids */
} poa_b200_final_graph_view_t;
int  poa_b200_block_final_graph(const poa_b200_block_view_t *view, int32_t padding_len, int32_t include_consensus,
                                poa_b200_graph_t **graph);
int  poa_b200_final_graph_view(const poa_b200_graph_t *graph, poa_b200_final_graph_view_t *view);

int64_t poa_b200_result_n_blocks(const poa_b200_result_t *res);
int  poa_b200_result_block(const poa_b200_result_t *res, int64_t block, poa_b200_block_view_t *view);
/* Drop the flat expansion of one block (views of it become invalid); a consumer that streams through a large batch calls it
 * after it has turned the block into its own graph. */
void poa_b200_result_release_block(const poa_b200_result_t *res, int64_t block);
/* Checksum of one finished block's graph: FNV-1a 64 over node_n (int32), then per node its base (one byte), out ids and
 * out weights (int32 each, final order) -- the same byte stream a checker can hash straight from abPOA's abpoa_graph_t
 * (node_n, node[i].base, node[i].out_id, node[i].out_edge_weight; deps/abPOA/include/abpoa.h:98-118), so that large samples
 * can be compared block by block without materialising either graph in the harness.  Pure host code. */
int  poa_b200_result_block_hash(const poa_b200_result_t *res, int64_t block, uint64_t *hash);
int  poa_b200_result_stats(const poa_b200_result_t *res, poa_b200_stats_t *stats);
void poa_b200_result_free(poa_b200_result_t *res);

#ifdef __cplusplus
}
#endif
#endif /* POA_B200_H */
