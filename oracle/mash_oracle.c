/* TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * Scalar C restatement of the per-block identity estimate smoothxg runs under --adaptive-poa-params
 * (src/smooth.cpp:1982-2062) and of the mkmh/rkmh routines it calls.  Pinned against the unmodified
 * reference headers (oracle/ref_mash_shim.cpp -> oracle/_ref/libmash_ref.so) by tests/test_mash_oracle.py
 * and against the golden vectors generated from them (tests/golden/mash_golden.npz).
 *
 *   murmur3_x64_128_h1   MurmurHash3_x64_128 (deps/mkmh/murmur3/murmur3.cpp:234-312), first 64-bit word only
 *   kmer_hash            mkmh::calc_hashes body (deps/mkmh/mkmh.hpp:512-534) for one position
 *   mash_hashes          calc_hashes(seq, len, k) (:768-774) + sort (rkmh.hpp:14-25)
 *   mash_common          the merge loop of rkmh::compare (rkmh.hpp:41-74)
 *   mash_distance        its distance formula (:76-93) with min_sketch_size_as_denom = true
 *   mash_block           src/smooth.cpp:1982-2023
 *   mash_preset          src/smooth.cpp:2026-2062
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

static inline uint64_t fmix64(uint64_t k) { /* murmur3.cpp:58-67 */
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

/* h1 of MurmurHash3_x64_128(key, len, seed): the only word calc_hashes reads (mkmh.hpp:527-528 use fhash[0], fhash[1],
 * the two 32-bit halves of h1). */
static uint64_t murmur3_x64_128_h1(const uint8_t *data, int len, uint32_t seed) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    const int nblocks = len / 16;
    for (int i = 0; i < nblocks; ++i) {
        uint64_t k1, k2;
        memcpy(&k1, data + 16 * i, 8); memcpy(&k2, data + 16 * i + 8, 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *tail = data + nblocks * 16;
    const int rem = len & 15;
    uint64_t k1 = 0, k2 = 0;
    for (int i = rem - 1; i >= 8; --i) k2 ^= (uint64_t)tail[i] << (8 * (i - 8));
    if (rem > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    for (int i = (rem > 8 ? 8 : rem) - 1; i >= 0; --i) k1 ^= (uint64_t)tail[i] << (8 * i);
    if (rem > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= (uint64_t)(int64_t)len; h2 ^= (uint64_t)(int64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

/* mkmh.hpp:146-160 valid_dna: only A C G T a c g t pass canonical() (:191-197).  Lower-case letters then index
 * rev_arr[] (:181-187, 26 entries for 'A'..'Z') out of bounds in reverse_complement (:213-224): undefined behaviour in
 * the reference, so lower-case input is outside the parity contract; here (and on the device) a k-mer holding anything
 * but upper-case ACGT hashes to 0, as a non-canonical k-mer does. */
static inline int comp_base(uint8_t c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 0; }

static uint64_t kmer_hash(const uint8_t *s, int k) {
    uint8_t rev[64];
    for (int i = 0; i < k; ++i) {
        const int c = comp_base(s[k - 1 - i]);
        if (!c) return 0; /* hashes[i] keeps its zero initialisation (mkmh.hpp:770) */
        rev[i] = (uint8_t)c;
    }
    const uint64_t f = murmur3_x64_128_h1(s, k, 42), r = murmur3_x64_128_h1(rev, k, 42);
    /* static_cast<uint64_t>(fhash[0]) << 32 | fhash[1]: the halves of h1 swapped (mkmh.hpp:527-528) */
    const uint64_t tf = (f << 32) | (f >> 32), tr = (r << 32) | (r >> 32);
    return tf < tr ? tf : tr;
}

static int cmp_u64(const void *a, const void *b) {
    const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

/* sorted hashes of one sequence; numhashes = len - k (mkmh.hpp:769: the last k-mer is not hashed); k <= 63 */
int mash_hashes(const char *seq, int len, int k, uint64_t *out) {
    const int n = len - k;
    for (int i = 0; i < n; ++i) out[i] = kmer_hash((const uint8_t *)seq + i, k);
    if (n > 0) qsort(out, (size_t)n, sizeof(uint64_t), cmp_u64);
    return n > 0 ? n : 0;
}

/* rkmh.hpp:41-74: number of merge matches between two sorted lists, zeros skipped; *denom = the union count it builds */
uint64_t mash_common(const uint64_t *a, int na, const uint64_t *b, int nb, uint64_t *denom) {
    int i = 0, j = 0;
    uint64_t common = 0, d;
    while (i < na && a[i] == 0) i++;
    while (j < nb && b[j] == 0) j++;
    d = (uint64_t)(i + j);
    while (i < na && j < nb) {
        if (a[i] == b[j]) { i++; j++; common++; }
        else if (a[i] > b[j]) j++;
        else i++;
        d++;
    }
    d += (uint64_t)(na - i); d += (uint64_t)(nb - j);
    if (denom) *denom = d;
    return common;
}

/* rkmh.hpp:76-93 with min_sketch_size_as_denom = true */
double mash_distance(uint64_t common, uint64_t denom, int na, int nb, int k) {
    const double jaccard = (double)common / (double)(na < nb ? na : nb);
    double distance;
    if (common == denom) distance = 0;
    else if (common == 0) distance = 1.;
    else {
        distance = -log(2 * jaccard / (1. + jaccard)) / k;
        if (distance > 1) distance = 1;
    }
    return distance;
}

static int cmp_f32(const void *a, const void *b) {
    const float x = *(const float *)a, y = *(const float *)b;
    return x < y ? -1 : x > y;
}

/* src/smooth.cpp:1982-2023.  Returns the number of sequences kept; with fewer than two, *threshold is untouched.
 * pair_identity / pair_common (optional): per (i, j>i) pair of kept sequences, in that order. */
int mash_block(int n_seq, const char *const *seq, const int *len, int kmer, float *threshold, float *pair_identity, uint64_t *pair_common) {
    int kept = 0;
    int *idx = (int *)malloc(sizeof(int) * (size_t)(n_seq > 0 ? n_seq : 1));
    for (int i = 0; i < n_seq; ++i)
        if ((size_t)len[i] >= (size_t)(8 * kmer)) idx[kept++] = i; /* :1996 */
    if (kept > 1) {
        uint64_t **h = (uint64_t **)malloc(sizeof(uint64_t *) * (size_t)kept);
        int *hn = (int *)malloc(sizeof(int) * (size_t)kept);
        for (int a = 0; a < kept; ++a) {
            h[a] = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)len[idx[a]]);
            hn[a] = mash_hashes(seq[idx[a]], len[idx[a]], kmer, h[a]);
        }
        const size_t np = (size_t)kept * (size_t)(kept - 1) / 2;
        float *est = (float *)malloc(sizeof(float) * np);
        size_t p = 0;
        for (int a = 0; a < kept; ++a)
            for (int b = a + 1; b < kept; ++b, ++p) {
                uint64_t denom;
                const uint64_t common = mash_common(h[a], hn[a], h[b], hn[b], &denom);
                est[p] = (float)(1.0 - mash_distance(common, denom, hn[a], hn[b], kmer)); /* :2014 */
                if (pair_common) pair_common[p] = common;
            }
        if (pair_identity) memcpy(pair_identity, est, sizeof(float) * np);
        qsort(est, np, sizeof(float), cmp_f32);
        const float q = est[(size_t)((double)(np - 1) * 0.30)]; /* :2021 */
        *threshold = q > 0.7f ? q : 0.7f;
        free(est);
        for (int a = 0; a < kept; ++a) free(h[a]);
        free(h); free(hn);
    }
    free(idx);
    return kept;
}

/* src/smooth.cpp:2026-2062: scores m,n,g,e,q,c for the estimated identity; returns 0 below 0.90 (scores left alone) */
int mash_preset(float t, int *s) {
    static const int tab[5][7] = {{99, 1, 19, 39, 3, 81, 1}, {98, 1, 13, 31, 3, 51, 1}, {97, 1, 9, 16, 2, 41, 1},
                                  {95, 1, 7, 11, 2, 33, 1}, {90, 1, 4, 6, 2, 26, 1}};
    static const double thr[5] = {0.99, 0.98, 0.97, 0.95, 0.90};
    for (int r = 0; r < 5; ++r)
        if (t >= thr[r]) { for (int c = 0; c < 6; ++c) s[c] = tab[r][c + 1]; return 1; }
    return 0;
}
