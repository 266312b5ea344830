// mash_b200.cu -- per-block identity estimate behind smoothxg's --adaptive-poa-params (include/mash_b200.h).
//
// Reference data flow (src/smooth.cpp:1982-2023): per block, hash every k-mer of every string (canonical
// MurmurHash3_x64_128 seed 42, deps/mkmh/mkmh.hpp:512-534), sort each list (deps/mkmh/rkmh.hpp:14-25), merge-compare
// all pairs (rkmh.hpp:41-96), 30th percentile of the identities.  Here, for a whole batch of blocks per call:
//
//   mash_hash_kernel     one CTA per string, the string staged through shared memory in 1 KB tiles; thread t hashes
//                        k-mer t of the tile (both strands) and writes one coalesced 8-byte word
//   mash_sort_kernel     one CTA per string: bitonic network in shared memory (lists up to 16 K hashes; longer ones run
//                        the same network in global memory).  All comparators point the same way (mirror step + half
//                        cleaners), so lists need no padding: a comparator whose upper index is past the end is a no-op
//   mash_compare_kernel  one CTA per block, one warp per pair: the merge loop of rkmh::compare counts, for every value,
//                        min(multiplicity in A, multiplicity in B) over the non-zero hashes; each lane takes elements of
//                        the shorter list, finds its occurrence number among equal neighbours and binary-searches the
//                        other list (which stays in L1/L2) -- no serial merge.  The union count the reference builds
//                        alongside is |A| + |B| - common (every element is consumed exactly once), so only `common`
//                        leaves the device.
//
// The floating-point tail (one libm log per pair, the percentile) runs on host threads with the reference's expressions.
// All three kernels are HBM/L2-bound integer work: 8 B written per k-mer by the hash, 16 B per hash in and out of the sort,
// and the compare re-reads lists that are L2-resident (a block's lists total S * L * 8 B, 0.5 MB for 32 x 2 kb).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mash_b200.h"
#include "mash_core.cuh"

namespace {

thread_local std::string g_mash_err;
int mash_fail(int code, const std::string &msg) { g_mash_err = msg; return code; }

#define MCU(call)                                                                                              \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess) { rc = mash_fail(MASH_B200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); goto done; } \
    } while (0)

constexpr int HASH_TILE = 1024;       // k-mer start positions per shared-memory tile
constexpr int HASH_THREADS = 256;
constexpr int SORT_THREADS = 512;
constexpr int SORT_SMEM_CAP = 16384;  // hashes sorted in shared memory (128 KB); longer lists sort in global memory
constexpr int CMP_THREADS = 512;

// mkmh::calc_hashes (mkmh.hpp:512-534,768-774): len - k hashes per string (the last k-mer is not hashed)
__global__ void __launch_bounds__(HASH_THREADS) mash_hash_kernel(const char *__restrict__ bases, const long long *__restrict__ src_off,
                                                                const int *__restrict__ len, const long long *__restrict__ hash_off,
                                                                int n_seq, int k, unsigned long long *__restrict__ hashes) {
    __shared__ uint8_t tile[HASH_TILE + 64];
    for (int s = blockIdx.x; s < n_seq; s += gridDim.x) {
        const char *g = bases + src_off[s];
        const int L = len[s], n = L - k;
        unsigned long long *out = hashes + hash_off[s];
        for (int p0 = 0; p0 < n; p0 += HASH_TILE) {
            const int nbytes = min(HASH_TILE + k - 1, L - p0);
            __syncthreads();
            for (int t = threadIdx.x; t < nbytes; t += HASH_THREADS) tile[t] = (uint8_t)g[p0 + t];
            __syncthreads();
            for (int t = threadIdx.x; t < HASH_TILE && p0 + t < n; t += HASH_THREADS) out[p0 + t] = mash::kmer_hash(tile + t, k);
        }
    }
}

// std::sort of each list (rkmh.hpp:20): ascending bitonic network, one CTA per list
__global__ void __launch_bounds__(SORT_THREADS) mash_sort_kernel(unsigned long long *__restrict__ hashes, const long long *__restrict__ hash_off,
                                                                const int *__restrict__ len, int n_seq, int k, int smem_cap) {
    extern __shared__ __align__(16) unsigned long long sm_sort[];
    for (int s = blockIdx.x; s < n_seq; s += gridDim.x) {
        const int n_ = len[s] - k;
        if (n_ <= 1) continue;  // uniform over the CTA
        const unsigned n = (unsigned)n_;
        unsigned long long *g = hashes + hash_off[s];
        const bool in_smem = n_ <= smem_cap;
        unsigned long long *a = in_smem ? sm_sort : g;
        __syncthreads();
        if (in_smem) for (unsigned t = threadIdx.x; t < n; t += SORT_THREADS) sm_sort[t] = g[t];
        __syncthreads();
        unsigned N = 2, logN = 1;
        while (N < n) { N <<= 1; ++logN; }
        for (unsigned ls = 1; ls <= logN; ++ls) {  // sorted runs of 2^ls
            for (unsigned t = threadIdx.x; t < (N >> 1); t += SORT_THREADS) mash::sort_mirror(a, n, ls, t);
            __syncthreads();
            for (unsigned lst = ls - 1; lst >= 1; --lst) {
                for (unsigned t = threadIdx.x; t < (N >> 1); t += SORT_THREADS) mash::sort_clean(a, n, lst, t);
                __syncthreads();
            }
        }
        if (in_smem) for (unsigned t = threadIdx.x; t < n; t += SORT_THREADS) g[t] = sm_sort[t];
    }
}

// merge-match count of rkmh::compare (rkmh.hpp:41-74) for every pair (i, j > i) of a block's kept strings
__global__ void __launch_bounds__(CMP_THREADS) mash_compare_kernel(const unsigned long long *__restrict__ hashes, const long long *__restrict__ hash_off,
                                                                  const int *__restrict__ len, const long long *__restrict__ blk_kept_off,
                                                                  const long long *__restrict__ pair_off, int n_blocks, int k,
                                                                  unsigned *__restrict__ pair_common) {
    const int warps = CMP_THREADS >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const long long s0 = blk_kept_off[b];
        const int kept = (int)(blk_kept_off[b + 1] - s0);
        if (kept < 2) continue;
        const long long np = (long long)kept * (kept - 1) / 2;
        for (long long p = wid; p < np; p += warps) {
            int i, j;
            mash::pair_decode(p, kept, i, j);
            const unsigned long long *A = hashes + hash_off[s0 + i], *B = hashes + hash_off[s0 + j];
            int na = len[s0 + i] - k, nb = len[s0 + j] - k;
            if (na > nb) { const unsigned long long *t = A; A = B; B = t; const int tn = na; na = nb; nb = tn; }  // probe with the shorter list
            unsigned cnt = 0;
            for (int q = lane; q < na; q += 32) cnt += mash::match_one(A, q, B, nb);
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0) pair_common[pair_off[b] + p] = cnt;
        }
    }
}

// rkmh.hpp:76-96 with min_sketch_size_as_denom = true, then src/smooth.cpp:2014
inline float est_identity(uint64_t common, uint64_t na, uint64_t nb, int kmer_size) {
    const uint64_t denom = na + nb - common;  // the union count the merge loop builds
    const double jaccard = double(common) / (double)std::min(na, nb);
    double distance;
    if (common == denom) distance = 0;
    else if (common == 0) distance = 1.;
    else {
        distance = -std::log(2 * jaccard / (1. + jaccard)) / kmer_size;
        if (distance > 1) distance = 1;
    }
    const float est = 1.0 - distance;
    return est;
}

template <class F>
void parallel_blocks(int64_t n, F &&f) {
    const unsigned hw = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
    const unsigned nt = (unsigned)std::min<int64_t>(hw, std::max<int64_t>(1, n / 64));
    if (nt <= 1) { for (int64_t b = 0; b < n; ++b) f(b); return; }
    std::atomic<int64_t> next{0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&] { for (;;) { const int64_t b0 = next.fetch_add(64); if (b0 >= n) break; for (int64_t b = b0; b < std::min(n, b0 + 64); ++b) f(b); } });
    for (auto &t : th) t.join();
}

}  // namespace

extern "C" {

const char *mash_b200_last_error(void) { return g_mash_err.c_str(); }

int64_t mash_b200_pair_offsets(int32_t kmer_size, int64_t n_blocks, const int64_t *block_seq_off, const int32_t *seq_len, int64_t *pair_off) {
    int64_t tot = 0;
    for (int64_t b = 0; b < n_blocks; ++b) {
        int64_t kept = 0;
        for (int64_t s = block_seq_off[b]; s < block_seq_off[b + 1]; ++s) kept += (int64_t)seq_len[s] >= 8LL * kmer_size;  // :1996
        if (pair_off) pair_off[b] = tot;
        if (kept > 1) tot += kept * (kept - 1) / 2;  // :2003
    }
    if (pair_off) pair_off[n_blocks] = tot;
    return tot;
}

int mash_b200_preset(float threshold, int32_t scores[6]) {  // src/smooth.cpp:2026-2062 (float compared with double literals)
    static const double thr[5] = {0.99, 0.98, 0.97, 0.95, 0.90};
    static const int32_t tab[5][6] = {{1, 19, 39, 3, 81, 1}, {1, 13, 31, 3, 51, 1}, {1, 9, 16, 2, 41, 1}, {1, 7, 11, 2, 33, 1}, {1, 4, 6, 2, 26, 1}};
    for (int r = 0; r < 5; ++r)
        if ((double)threshold >= thr[r]) { for (int c = 0; c < 6; ++c) scores[c] = tab[r][c]; return 1; }
    return 0;
}

int mash_b200_block_identity(int device, int32_t kmer_size, int64_t n_blocks, const int64_t *block_seq_off, const int32_t *seq_len,
                             const int64_t *seq_off, const char *bases, float *threshold, int32_t *n_kept, uint32_t *pair_common,
                             float *pair_identity, mash_b200_stats_t *stats) {
    if (kmer_size < 1 || kmer_size > 32) return mash_fail(MASH_B200_EARG, "kmer_size must be 1..32");
    if (n_blocks < 0 || (n_blocks > 0 && (!block_seq_off || !seq_len || !seq_off || !bases || !threshold))) return mash_fail(MASH_B200_EARG, "null argument");
    mash_b200_stats_t st{};
    if (n_blocks == 0) { if (stats) *stats = st; return MASH_B200_OK; }
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) { cudaGetLastError(); return mash_fail(MASH_B200_ECUDA, "no usable CUDA device (there is no CPU path)"); }
    if (cudaSetDevice(device) != cudaSuccess) return mash_fail(MASH_B200_ECUDA, "cudaSetDevice failed");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return mash_fail(MASH_B200_ECUDA, "cudaGetDeviceProperties failed");

    // ---- kept strings (:1996), their hash-list and pair offsets
    std::vector<int64_t> blk_kept_off(n_blocks + 1, 0), pair_off(n_blocks + 1, 0);
    std::vector<int64_t> kept_src;  // byte offset of each kept string in `bases`
    std::vector<int32_t> kept_len;
    for (int64_t b = 0; b < n_blocks; ++b) {
        if (block_seq_off[b + 1] < block_seq_off[b]) return mash_fail(MASH_B200_EARG, "block_seq_off not monotonic");
        for (int64_t s = block_seq_off[b]; s < block_seq_off[b + 1]; ++s) {
            if (seq_len[s] < 0 || seq_off[s + 1] < seq_off[s] || seq_off[s + 1] - seq_off[s] < seq_len[s]) return mash_fail(MASH_B200_EARG, "seq_off/seq_len inconsistent");
            if ((int64_t)seq_len[s] >= 8LL * kmer_size) { kept_src.push_back(seq_off[s]); kept_len.push_back(seq_len[s]); }
        }
        blk_kept_off[b + 1] = (int64_t)kept_src.size();
        const int64_t kept = blk_kept_off[b + 1] - blk_kept_off[b];
        pair_off[b + 1] = pair_off[b] + (kept > 1 ? kept * (kept - 1) / 2 : 0);
        if (n_kept) n_kept[b] = (int32_t)kept;
    }
    std::vector<uint32_t> common_host((size_t)pair_off[n_blocks]);
    st.n_seqs_kept = (int64_t)kept_src.size();
    st.n_pairs = pair_off[n_blocks];

    // ---- device passes over chunks of blocks (bounded device memory: <= 4 GB of hashes per chunk)
    int64_t chunk_hashes = 1LL << 29;
    if (const char *e = getenv("MASH_B200_CHUNK_HASHES")) { const long long v = atoll(e); if (v > 0) chunk_hashes = v; }  // tests: force several chunks
    int rc = MASH_B200_OK;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    char *d_bases = nullptr; long long *d_src = nullptr, *d_hoff = nullptr, *d_bko = nullptr, *d_poff = nullptr; int *d_len = nullptr;
    unsigned long long *d_hash = nullptr; unsigned *d_common = nullptr;
    size_t cap_bases = 0, cap_seqs = 0, cap_blocks = 0, cap_hash = 0, cap_pairs = 0;
    auto release = [&] {
        cudaFree(d_bases); cudaFree(d_src); cudaFree(d_hoff); cudaFree(d_bko); cudaFree(d_poff); cudaFree(d_len); cudaFree(d_hash); cudaFree(d_common);
        d_bases = nullptr; d_src = d_hoff = d_bko = d_poff = nullptr; d_len = nullptr; d_hash = nullptr; d_common = nullptr;
    };
    {
        MCU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        for (auto &e : ev) MCU(cudaEventCreate(&e));
        MCU(cudaFuncSetAttribute(mash_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_SMEM_CAP * 8));
        int64_t b0 = 0;
        while (b0 < n_blocks) {
            // chunk [b0, b1): whole blocks, at least one
            int64_t b1 = b0, nh = 0;
            while (b1 < n_blocks) {
                int64_t add = 0;
                for (int64_t s = blk_kept_off[b1]; s < blk_kept_off[b1 + 1]; ++s) add += kept_len[s] - kmer_size;
                if (b1 > b0 && nh + add > chunk_hashes) break;
                nh += add; ++b1;
            }
            const int64_t s_lo = blk_kept_off[b0], s_hi = blk_kept_off[b1], ns = s_hi - s_lo, nb = b1 - b0;
            const int64_t np = pair_off[b1] - pair_off[b0];
            st.n_chunks++;
            if (ns > 0 && np > 0) {
                // byte range of the chunk's kept strings
                int64_t byte_lo = kept_src[s_lo], byte_hi = 0;
                for (int64_t s = s_lo; s < s_hi; ++s) { byte_lo = std::min(byte_lo, kept_src[s]); byte_hi = std::max(byte_hi, kept_src[s] + kept_len[s]); }
                std::vector<long long> h_src(ns), h_hoff(ns + 1), h_bko(nb + 1), h_poff(nb + 1);
                int max_n = 0;
                h_hoff[0] = 0;
                for (int64_t s = 0; s < ns; ++s) {
                    h_src[s] = kept_src[s_lo + s] - byte_lo;
                    const int n = kept_len[s_lo + s] - kmer_size;
                    h_hoff[s + 1] = h_hoff[s] + n;
                    max_n = std::max(max_n, n);
                }
                for (int64_t b = 0; b <= nb; ++b) { h_bko[b] = blk_kept_off[b0 + b] - s_lo; h_poff[b] = pair_off[b0 + b] - pair_off[b0]; }
                const size_t nbytes = (size_t)(byte_hi - byte_lo);
                if (nbytes > cap_bases || (size_t)ns > cap_seqs || (size_t)nb > cap_blocks || (size_t)nh > cap_hash || (size_t)np > cap_pairs) {
                    release();
                    cap_bases = nbytes; cap_seqs = (size_t)ns; cap_blocks = (size_t)nb; cap_hash = (size_t)nh; cap_pairs = (size_t)np;
                    MCU(cudaMalloc(&d_bases, cap_bases)); MCU(cudaMalloc(&d_src, cap_seqs * 8)); MCU(cudaMalloc(&d_hoff, (cap_seqs + 1) * 8));
                    MCU(cudaMalloc(&d_len, cap_seqs * 4)); MCU(cudaMalloc(&d_bko, (cap_blocks + 1) * 8)); MCU(cudaMalloc(&d_poff, (cap_blocks + 1) * 8));
                    MCU(cudaMalloc(&d_hash, std::max<size_t>(cap_hash, 1) * 8)); MCU(cudaMalloc(&d_common, cap_pairs * 4));
                }
                MCU(cudaEventRecord(ev[0], stream));
                MCU(cudaMemcpyAsync(d_bases, bases + byte_lo, nbytes, cudaMemcpyHostToDevice, stream));
                MCU(cudaMemcpyAsync(d_src, h_src.data(), (size_t)ns * 8, cudaMemcpyHostToDevice, stream));
                MCU(cudaMemcpyAsync(d_hoff, h_hoff.data(), (size_t)(ns + 1) * 8, cudaMemcpyHostToDevice, stream));
                MCU(cudaMemcpyAsync(d_len, kept_len.data() + s_lo, (size_t)ns * 4, cudaMemcpyHostToDevice, stream));
                MCU(cudaMemcpyAsync(d_bko, h_bko.data(), (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, stream));
                MCU(cudaMemcpyAsync(d_poff, h_poff.data(), (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, stream));
                MCU(cudaEventRecord(ev[1], stream));
                const int sms = prop.multiProcessorCount;
                const int grid_h = (int)std::min<int64_t>(ns, (int64_t)sms * 32);
                mash_hash_kernel<<<grid_h, HASH_THREADS, 0, stream>>>(d_bases, d_src, d_len, d_hoff, (int)ns, kmer_size, d_hash);
                MCU(cudaGetLastError());
                MCU(cudaEventRecord(ev[2], stream));
                const int smem_elems = std::min(max_n, SORT_SMEM_CAP);
                const int grid_s = (int)std::min<int64_t>(ns, (int64_t)sms * 16);
                mash_sort_kernel<<<grid_s, SORT_THREADS, (size_t)std::max(smem_elems, 1) * 8, stream>>>(d_hash, d_hoff, d_len, (int)ns, kmer_size, smem_elems);
                MCU(cudaGetLastError());
                MCU(cudaEventRecord(ev[3], stream));
                const int grid_c = (int)std::min<int64_t>(nb, (int64_t)sms * 4);
                mash_compare_kernel<<<grid_c, CMP_THREADS, 0, stream>>>(d_hash, d_hoff, d_len, d_bko, d_poff, (int)nb, kmer_size, d_common);
                MCU(cudaGetLastError());
                MCU(cudaEventRecord(ev[4], stream));
                MCU(cudaMemcpyAsync(common_host.data() + pair_off[b0], d_common, (size_t)np * 4, cudaMemcpyDeviceToHost, stream));
                MCU(cudaEventRecord(ev[5], stream));
                MCU(cudaStreamSynchronize(stream));
                float ms;
                MCU(cudaEventElapsedTime(&ms, ev[0], ev[1])); st.h2d_ms += ms;
                MCU(cudaEventElapsedTime(&ms, ev[1], ev[2])); st.hash_ms += ms;
                MCU(cudaEventElapsedTime(&ms, ev[2], ev[3])); st.sort_ms += ms;
                MCU(cudaEventElapsedTime(&ms, ev[3], ev[4])); st.compare_ms += ms;
                MCU(cudaEventElapsedTime(&ms, ev[4], ev[5])); st.d2h_ms += ms;
                st.kernel_launches += 3;
                st.n_hashes += nh;
                st.h2d_bytes += (int64_t)nbytes + ns * 20 + 8 + (nb + 1) * 16;
                st.d2h_bytes += np * 4;
            }
            b0 = b1;
        }
    }
    {
        // ---- floating-point tail on host threads (rkmh.hpp:76-96, src/smooth.cpp:2014-2021)
        const auto t0 = std::chrono::steady_clock::now();
        parallel_blocks(n_blocks, [&](int64_t b) {
            const int64_t s0 = blk_kept_off[b], kept = blk_kept_off[b + 1] - s0;
            if (kept < 2) { threshold[b] = -1.0f; return; }
            const int64_t np = kept * (kept - 1) / 2, po = pair_off[b];
            std::vector<float> est((size_t)np);
            int64_t p = 0;
            for (int64_t i = 0; i < kept; ++i)
                for (int64_t j = i + 1; j < kept; ++j, ++p)
                    est[(size_t)p] = est_identity(common_host[(size_t)(po + p)], (uint64_t)(kept_len[s0 + i] - kmer_size), (uint64_t)(kept_len[s0 + j] - kmer_size), kmer_size);
            if (pair_identity) std::copy(est.begin(), est.end(), pair_identity + po);
            std::sort(est.begin(), est.end());
            threshold[b] = std::max((float)0.7, est[(est.size() - 1) * 0.30]);
        });
        if (pair_common && !common_host.empty()) std::memcpy(pair_common, common_host.data(), common_host.size() * 4);
        st.host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
done:
    release();
    for (auto &e : ev) if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
    if (stats) *stats = st;
    return rc;
}

}  // extern "C"
