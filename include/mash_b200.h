/* mash_b200.h -- C ABI of the per-block identity estimate behind smoothxg's --adaptive-poa-params
 * (the step in front of the POA hot path).
 *
 * Reference: smooth_and_lace, src/smooth.cpp:1982-2062.  For every block with 2..max_block_depth path ranges it
 *   (1) concatenates each range's node sequences (:1987-1992; needs the XG index, stays in smoothxg),
 *   (2) drops strings shorter than 8*kmer_size (:1996-2000),
 *   (3) hashes every k-mer of every string -- canonical MurmurHash3_x64_128, seed 42 -- and sorts each list
 *       (rkmh::hash_sequences, deps/mkmh/rkmh.hpp:14-25; mkmh::calc_hashes, deps/mkmh/mkmh.hpp:512-534,768-774),
 *   (4) compares all pairs (rkmh::compare, rkmh.hpp:41-96, min sketch size as denominator),
 *   (5) takes the 30th percentile of the estimated identities, clamped at 0.7 (:2020-2021), and
 *   (6) picks one of five score presets from it (:2026-2062).
 * Steps (2)-(5) are what this ABI replaces, for a whole batch of blocks per call: (3) and the integer part of (4)
 * -- O(S*L) hashes, O(S*L*log L) sorting and O(S^2 * L) list intersection per block -- run in hand-written sm_100a
 * kernels (smoothxg_b200/csrc/mash_b200.cu); the O(S^2) floating-point tail of (4) (one libm log per pair) and (5)
 * run on host threads with the reference's own expressions, so thresholds are bit-identical to the reference's.
 * (6) is mash_b200_preset() (pure host).  There is no CPU fallback for (3)/(4): without a usable device the call fails.
 *
 * Input strings are upper-case ASCII as XG stores them.  K-mers holding anything but A, C, G, T hash to 0 and are
 * skipped by the comparison, as in the reference (mkmh.hpp:191-197, rkmh.hpp:48-53).  Lower-case a/c/g/t are
 * outside the contract: the reference indexes a 26-entry table out of bounds for them (mkmh.hpp:181-224).
 */
#ifndef MASH_B200_H
#define MASH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    MASH_B200_OK    = 0,
    MASH_B200_ECUDA = 6, /* CUDA runtime error; see mash_b200_last_error() (same numbering as poa_b200.h) */
    MASH_B200_EARG  = 7,
    MASH_B200_ENOMEM = 8
};

typedef struct mash_b200_stats {
    double  h2d_ms, hash_ms, sort_ms, compare_ms, d2h_ms, host_ms; /* device phases by CUDA events on the launch stream */
    int64_t n_seqs_kept, n_hashes, n_pairs;
    int64_t h2d_bytes, d2h_bytes;
    int32_t kernel_launches, n_chunks;
} mash_b200_stats_t;

const char *mash_b200_last_error(void);

/* Pairs a block contributes: kept*(kept-1)/2 with kept = sequences of at least 8*kmer_size bases (0 if kept < 2).
 * Fills pair_off[n_blocks+1] (prefix sums); returns the total.  Pure host code. */
int64_t mash_b200_pair_offsets(int32_t kmer_size, int64_t n_blocks, const int64_t *block_seq_off, const int32_t *seq_len,
                               int64_t *pair_off);

/* Steps (2)-(5) for a batch of blocks.
 *   block_seq_off[n_blocks+1] -> index into seq_len; seq_off[n_seqs+1] -> byte offset of each string in `bases`.
 *   threshold[n_blocks]: est_identity_threshold (:2021); -1 for blocks where fewer than two strings qualify (the
 *     reference then keeps the user's scores, :2003).  n_kept[n_blocks] (optional): strings that qualified.
 *   pair_common / pair_identity (optional, mash_b200_pair_offsets() layout, (i, j>i) order over kept strings):
 *     the merge-match count of rkmh::compare and the estimated identity (:2014) of every pair.
 * kmer_size: 1..32 (smoothxg's default is 17, src/main.cpp:304). */
int mash_b200_block_identity(int device, int32_t kmer_size, int64_t n_blocks, const int64_t *block_seq_off,
                             const int32_t *seq_len, const int64_t *seq_off, const char *bases,
                             float *threshold, int32_t *n_kept, uint32_t *pair_common, float *pair_identity,
                             mash_b200_stats_t *stats);

/* Step (6), src/smooth.cpp:2026-2062: scores[6] = poa_m, poa_n, poa_g, poa_e, poa_q, poa_c for the threshold.
 * Returns 1 and fills scores, or 0 (threshold < 0.90, or -1: keep the user's scores). */
int mash_b200_preset(float threshold, int32_t scores[6]);

#ifdef __cplusplus
}
#endif
#endif /* MASH_B200_H */
