// Host replay of the device logic of smoothxg_b200/csrc/mash_core.cuh (TEST INFRASTRUCTURE): the same functions the
// kernels of mash_b200.cu call, compiled for the CPU (MASH_HOST_EMU), with the kernels' loops over threads run serially
// (every loop between two __syncthreads() is a loop over t here).  tests/test_mash_emu.py compares the output with the
// oracle restatement and the unmodified reference headers.
#define MASH_HOST_EMU
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../smoothxg_b200/csrc/mash_core.cuh"

extern "C" {

// mash_hash_kernel + mash_sort_kernel for one string: returns len - k sorted hashes
int emu_mash_hashes(const char *seq, int len, int k, uint64_t *out) {
    const int n = len - k;
    if (n <= 0) return 0;
    std::vector<unsigned long long> a((size_t)n);
    for (int t = 0; t < n; ++t) a[(size_t)t] = mash::kmer_hash((const uint8_t *)seq + t, k);
    if (n > 1) {
        unsigned N = 2, logN = 1;
        while (N < (unsigned)n) { N <<= 1; ++logN; }
        for (unsigned ls = 1; ls <= logN; ++ls) {
            for (unsigned t = 0; t < (N >> 1); ++t) mash::sort_mirror(a.data(), (unsigned)n, ls, t);
            for (unsigned lst = ls - 1; lst >= 1; --lst)
                for (unsigned t = 0; t < (N >> 1); ++t) mash::sort_clean(a.data(), (unsigned)n, lst, t);
        }
    }
    for (int t = 0; t < n; ++t) out[t] = a[(size_t)t];
    return n;
}

// mash_compare_kernel for one block of already-kept strings: pair_common[kept*(kept-1)/2]
void emu_mash_block_common(int kept, const char *const *seq, const int *len, int k, uint32_t *pair_common) {
    std::vector<std::vector<uint64_t>> h((size_t)kept);
    for (int s = 0; s < kept; ++s) { h[(size_t)s].resize((size_t)(len[s] > 0 ? len[s] : 1)); h[(size_t)s].resize((size_t)emu_mash_hashes(seq[s], len[s], k, h[(size_t)s].data())); }
    const long long np = (long long)kept * (kept - 1) / 2;
    for (long long p = 0; p < np; ++p) {
        int i, j;
        mash::pair_decode(p, kept, i, j);
        const unsigned long long *A = (const unsigned long long *)h[(size_t)i].data(), *B = (const unsigned long long *)h[(size_t)j].data();
        int na = (int)h[(size_t)i].size(), nb = (int)h[(size_t)j].size();
        if (na > nb) { const unsigned long long *t = A; A = B; B = t; const int tn = na; na = nb; nb = tn; }
        unsigned cnt = 0;
        for (int lane = 0; lane < 32; ++lane)
            for (int q = lane; q < na; q += 32) cnt += mash::match_one(A, q, B, nb);
        pair_common[p] = cnt;
    }
}
}
