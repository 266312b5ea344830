// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Thin shim over the UNMODIFIED mkmh / rkmh headers vendored by smoothxg (deps/mkmh/mkmh.hpp, rkmh.hpp,
// murmur3/murmur3.cpp), compiled from where they lie under /root/reference (oracle/Makefile).  It replays the
// per-block identity estimate smooth_and_lace runs when --adaptive-poa-params is given
// (src/smooth.cpp:1982-2023): keep sequences of at least 8*kmer bases, rkmh::hash_sequences, all-vs-all
// rkmh::compare(.., kmer, true), sort, 30th percentile clamped at 0.7.  The XG walk that produces the
// strings (:1987-1992) is replaced by the caller handing the strings in.
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "rkmh.hpp"

extern "C" {

// Returns the number of sequences kept (>= 8*kmer bases).  When at least two are kept, *threshold receives
// est_identity_threshold (:2021) and pair_identity (if not NULL, room for kept*(kept-1)/2 floats) the
// estimated identities in (i, j>i) order before sorting (:2013-2016); otherwise *threshold is left alone.
int mash_ref_block(int n_seq, const char *const *seq, const int *len, int kmer, float *threshold, float *pair_identity) {
    std::vector<std::string *> seqs;
    for (int i = 0; i < n_seq; ++i) {
        auto s = new std::string(seq[i], (size_t)len[i]);
        if (s->size() >= (size_t)(8 * kmer)) seqs.push_back(s); else delete s;
    }
    const int kept = (int)seqs.size();
    if (seqs.size() > 1) {
        std::vector<std::vector<mkmh::hash_t>> seq_hashes(seqs.size());
        std::vector<int> seq_hash_lens(seqs.size());
        rkmh::hash_sequences(seqs, seq_hashes, seq_hash_lens, kmer);
        std::vector<float> est;
        est.reserve(seqs.size() * (seqs.size() - 1) / 2);
        for (uint64_t i = 0; i < seqs.size(); ++i)
            for (uint64_t j = i + 1; j < seqs.size(); ++j) {
                const float est_identity = 1.0 - rkmh::compare(seq_hashes[i], seq_hashes[j], kmer, true);
                est.push_back(est_identity);
            }
        if (pair_identity) std::copy(est.begin(), est.end(), pair_identity);
        std::sort(est.begin(), est.end());
        *threshold = std::max((float)0.7, est[(est.size() - 1) * 0.30]);
    }
    for (auto &s : seqs) delete s;
    return kept;
}

// sorted hash list of one sequence (rkmh::hash_sequence, deps/mkmh/rkmh.hpp:27-31); returns its length (len - kmer)
int mash_ref_hashes(const char *seq, int len, int kmer, uint64_t *out) {
    std::vector<mkmh::hash_t> h = rkmh::hash_sequence(seq, len, kmer);
    std::copy(h.begin(), h.end(), out);
    return (int)h.size();
}
}
