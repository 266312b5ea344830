// mash_core.cuh -- device logic of the identity estimate (mash_b200.cu), written so that tests/emu/emu_mash.cpp can
// compile the same functions for the host (MASH_HOST_EMU) and replay the kernels' index arithmetic serially.
#pragma once
#include <stdint.h>

#ifdef MASH_HOST_EMU
#define MASH_D static inline
#else
#define MASH_D __device__ __forceinline__
#endif

namespace mash {

MASH_D uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
MASH_D uint64_t fmix64(uint64_t k) {  // murmur3.cpp:58-67
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}
MASH_D uint8_t comp_base(uint8_t c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A'; }

// little-endian 8-byte word w of the k-mer at s (RC: of its reverse complement), zero past the k-mer's end
template <bool RC>
MASH_D uint64_t kmer_word(const uint8_t *s, int k, int w) {
    uint64_t v = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const int i = 8 * w + b;
        if (i < k) v |= (uint64_t)(RC ? comp_base(s[k - 1 - i]) : s[i]) << (8 * b);
    }
    return v;
}

// h1 of MurmurHash3_x64_128(kmer, k, 42) (murmur3.cpp:234-312); the zero-padded tail words equal the switch's xors
template <bool RC>
MASH_D uint64_t murmur3_h1(const uint8_t *s, int k) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = 42, h2 = 42;
    const int nblocks = k >> 4;
    for (int i = 0; i < nblocks; ++i) {
        uint64_t k1 = kmer_word<RC>(s, k, 2 * i), k2 = kmer_word<RC>(s, k, 2 * i + 1);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const int rem = k & 15;
    if (rem > 8) { uint64_t k2 = kmer_word<RC>(s, k, 2 * nblocks + 1); k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    if (rem > 0) { uint64_t k1 = kmer_word<RC>(s, k, 2 * nblocks); k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= (uint64_t)k; h2 ^= (uint64_t)k;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    return h1 + h2;
}

// one entry of mkmh::calc_hashes (mkmh.hpp:512-534): 0 for a k-mer with a non-ACGT byte, else the minimum over strands of
// h1 with its 32-bit halves swapped (static_cast<uint64_t>(fhash[0]) << 32 | fhash[1])
MASH_D uint64_t kmer_hash(const uint8_t *sp, int k) {
    bool ok = true;
    for (int i = 0; i < k; ++i) { const uint8_t c = sp[i]; ok &= (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T'); }
    if (!ok) return 0;
    const uint64_t f = murmur3_h1<false>(sp, k), r = murmur3_h1<true>(sp, k);
    const uint64_t tf = (f << 32) | (f >> 32), tr = (r << 32) | (r >> 32);
    return tf < tr ? tf : tr;
}

MASH_D void cmp_exchange(unsigned long long *a, unsigned i, unsigned j) {
    const unsigned long long x = a[i], y = a[j];
    if (x > y) { a[i] = y; a[j] = x; }
}
// Ascending bitonic network whose comparators all point the same way, so a list of n < 2^logN entries needs no padding
// (a comparator whose upper index is >= n would compare with +infinity: a no-op).  Comparator t (0 <= t < 2^(logN-1)) of
// the mirror step that merges sorted runs of 2^(ls-1) into runs of 2^ls ...
MASH_D void sort_mirror(unsigned long long *a, unsigned n, unsigned ls, unsigned t) {
    const unsigned size = 1u << ls, half = size >> 1;
    const unsigned base = (t >> (ls - 1)) << ls, off = t & (half - 1);
    const unsigned i = base + off, j = base + size - 1 - off;
    if (j < n) cmp_exchange(a, i, j);
}
// ... and of the half cleaner with stride 2^(lst-1) that follows it (lst = ls-1 .. 1)
MASH_D void sort_clean(unsigned long long *a, unsigned n, unsigned lst, unsigned t) {
    const unsigned stride = 1u << (lst - 1);
    const unsigned i = ((t >> (lst - 1)) << lst) + (t & (stride - 1)), j = i + stride;
    if (j < n) cmp_exchange(a, i, j);
}

// pair number p (row-major over i < j) of `kept` strings
MASH_D void pair_decode(long long p, int kept, int &i, int &j) {
    i = 0;
    while (p >= kept - 1 - i) { p -= kept - 1 - i; ++i; }
    j = i + 1 + (int)p;
}

// Contribution of element q of sorted list A to the merge-match count of rkmh::compare (rkmh.hpp:41-74) against sorted
// list B: the merge matches min(multiplicity in A, multiplicity in B) copies of every non-zero value, i.e. copy number
// `occ` of a value in A is matched iff B holds more than occ copies of it.
MASH_D unsigned match_one(const unsigned long long *A, int q, const unsigned long long *B, int nb) {
    const unsigned long long v = A[q];
    if (v == 0) return 0;  // leading zeros are skipped (rkmh.hpp:48-53)
    int occ = 0;
    while (q - 1 - occ >= 0 && A[q - 1 - occ] == v) ++occ;
    int lo = 0, hi = nb;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (B[mid] < v) lo = mid + 1; else hi = mid; }
    return (lo + occ < nb && B[lo + occ] == v) ? 1u : 0u;
}

}  // namespace mash
