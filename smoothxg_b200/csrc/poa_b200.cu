// poa_b200.cu -- host side of the C ABI (include/poa_b200.h) and the kernel entry points.
//
// Host responsibilities (the counterpart of what smooth_abpoa does around abpoa_poa, reference
// src/smooth.cpp:256-351): validate parameters, build the 5x5 score matrix (abpoa_align.c:12-25),
// order blocks by cost, size and carve per-CTA workspaces out of HBM, launch one persistent kernel
// per batch, re-run blocks that exhausted a workspace with larger ones, and expose the flat result.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/poa_b200.h"
#include "poa_core.cuh"
#include "poa_host.hpp"
#include "poa_wire.hpp"
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <thread>
#include <unordered_map>

using namespace poa;

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
#ifndef POA_MIN_BLOCKS
// Resident single-warp POA blocks per SM the register budget is sized for.  ptxas maps any hint of 13..16 to 128 registers
// per thread (16 resident blocks); the hint still steers its register allocation.  Round 1 (before the row-loop diet): 12 x 168
// registers 217.9, 13 x 152 214.0, 14 x 144 208.7, 15 x 136 206.4, 16 x 128 221.1 Gcells/s on configs[2], and 13 gave the leanest
// chunk loop of the 128-register builds.  Re-measured on the final round-2 code (same box, scripts/lab_r02/gpu_r02_z.sh): 13 -> 273.3,
// 14 -> 276.8, 15 -> 274.9 Gcells/s; 14 also spills least (24 bytes, touched once per row at most, against 112).
#define POA_MIN_BLOCKS 14
#endif
// POA_MAXNREG (optional): cap registers per thread directly instead of through the resident-block hint
#ifdef POA_MAXNREG
#define POA_KERNEL_BOUNDS __maxnreg__(POA_MAXNREG)
#else
#define POA_KERNEL_BOUNDS __launch_bounds__(NW * 32, POA_MIN_BLOCKS / NW)
#endif
template <int NW>
__global__ void POA_KERNEL_BOUNDS poa_b200_block_kernel(const __grid_constant__ DevParams P, const __grid_constant__ DevBatch B, const __grid_constant__ WsLayout L, char *ws_base, const __grid_constant__ DevOut O) {
    __shared__ Shared sh;
    extern __shared__ __align__(16) char dyn_smem[];  // NW == 1: P16_SMEM_BYTES, else p16_mw_smem_bytes<NW>()
    constexpr int dyn_bytes = NW == 1 ? P16_SMEM_BYTES : p16_mw_smem<NW>::bytes;
    if (threadIdx.x == 0) { ws_bind(sh.ws, ws_base + (long long)blockIdx.x * L.stride, L); sh.ring = dyn_smem; sh.ring_bytes = dyn_bytes; }
    for (;;) {
        if (threadIdx.x == 0) sh.blk = atomicAdd(O.counter, 1);
        __syncthreads();
        const int t = sh.blk;
        __syncthreads();
        if (t >= B.n_order) break;
        poa_block<NW>(sh, P, B, L, O, B.order[t], ws_base + (long long)blockIdx.x * L.stride);
    }
}

// Base codes index the 5x5 score matrix on the device: anything above 4 (the reference's own encoder cannot produce it,
// abpoa_seq.c:15-32) is clamped to 4 = N once, right after the upload, instead of being trusted in every kernel.
__global__ void poa_b200_sanitize_bases_kernel(uint8_t *bases, long long n) {
    const long long n16 = n / 16;
    uint4 *v = reinterpret_cast<uint4 *>(bases);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
        uint4 u = v[i];
        const uint4 o = u;
        u.x = __vminu4(u.x, 0x04040404u); u.y = __vminu4(u.y, 0x04040404u); u.z = __vminu4(u.z, 0x04040404u); u.w = __vminu4(u.w, 0x04040404u);
        if (u.x != o.x || u.y != o.y || u.z != o.z || u.w != o.w) v[i] = u;
    }
    for (long long i = n16 * 16 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        if (bases[i] > 4) bases[i] = 4;
}

namespace poa {
thread_local std::string g_last_error;
int set_err(int code, const std::string &msg) { g_last_error = msg; return code; }
}  // namespace poa

namespace {

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return set_err(POA_B200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
    } while (0)

// Pinned host buffers are expensive to create (page-locking), so finished results hand theirs back
// to a pool shared with the engine instead of freeing them.
struct PinnedPool {
    std::mutex mu;
    std::vector<std::pair<void *, size_t>> free_list;
    void *take(size_t bytes, size_t *cap) {
        {
            std::lock_guard<std::mutex> lk(mu);
            size_t best = free_list.size();
            for (size_t i = 0; i < free_list.size(); ++i)
                if (free_list[i].second >= bytes && (best == free_list.size() || free_list[i].second < free_list[best].second)) best = i;
            if (best != free_list.size()) {
                void *p = free_list[best].first; *cap = free_list[best].second;
                free_list.erase(free_list.begin() + (long)best);
                return p;
            }
        }
        void *p = nullptr;
        size_t want = bytes + bytes / 8 + 4096;
        if (cudaMallocHost(&p, want) != cudaSuccess) {
            // page-locking can fail where pageable memory is still plentiful (locked-memory limits, many ranks on one host):
            // drop the pooled buffers and retry, then fall back to ordinary host memory (the copy is staged by the driver)
            cudaGetLastError();
            trim();
            if (cudaMallocHost(&p, want) != cudaSuccess) {
                cudaGetLastError();
                p = malloc(want);
                if (!p) return nullptr;
                std::lock_guard<std::mutex> lk(mu);
                pageable.push_back(p);
            }
        }
        *cap = want;
        return p;
    }
    void give(void *p, size_t cap) { std::lock_guard<std::mutex> lk(mu); free_list.emplace_back(p, cap); }
    void release(void *p) {
        auto it = std::find(pageable.begin(), pageable.end(), p);
        if (it != pageable.end()) { pageable.erase(it); free(p); } else cudaFreeHost(p);
    }
    void trim() { std::lock_guard<std::mutex> lk(mu); for (auto &e : free_list) release(e.first); free_list.clear(); }
    ~PinnedPool() { for (auto &e : free_list) release(e.first); }
    std::vector<void *> pageable;  // buffers that came from malloc()
};

// Same idea for device memory: cudaMalloc/cudaFree of tens of GB per batch would sit inside every
// end-to-end call, so workspaces and arenas are recycled through the engine.
struct DevicePool {
    std::mutex mu;
    std::vector<std::pair<void *, size_t>> free_list;
    void *take(size_t bytes, size_t *cap) {
        {
            std::lock_guard<std::mutex> lk(mu);
            size_t best = free_list.size();
            for (size_t i = 0; i < free_list.size(); ++i)
                if (free_list[i].second >= bytes && (best == free_list.size() || free_list[i].second < free_list[best].second)) best = i;
            if (best != free_list.size() && free_list[best].second <= 2 * bytes + (64u << 20)) {
                void *p = free_list[best].first; *cap = free_list[best].second;
                free_list.erase(free_list.begin() + (long)best);
                return p;
            }
        }
        void *p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            cudaGetLastError();
            trim();
            if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        }
        *cap = bytes;
        return p;
    }
    void give(void *p, size_t cap) { if (p) { std::lock_guard<std::mutex> lk(mu); free_list.emplace_back(p, cap); } }
    void trim() { std::lock_guard<std::mutex> lk(mu); for (auto &e : free_list) cudaFree(e.first); free_list.clear(); }
    size_t pooled_bytes() { std::lock_guard<std::mutex> lk(mu); size_t s = 0; for (auto &e : free_list) s += e.second; return s; }
    ~DevicePool() { for (auto &e : free_list) cudaFree(e.first); }
};

struct Arena {
    size_t cap_bytes = 0;
    int *d = nullptr;
    unsigned long long cap = 0;   // words
    unsigned long long used = 0;  // words, filled after the launch
};

}  // namespace

// Coalescing front end of the per-block entry points (poa_b200_submit_block / poa_b200_wait_block / poa_b200_poa_block):
// blocks submitted by any number of host threads accumulate in one pending batch; a dispatcher thread owned by the engine
// closes the batch and runs it through poa_b200_run_batch when it is large enough, or as soon as some thread waits for one of
// its blocks while the GPU is idle.  While a batch runs, further submissions accumulate for the next one, so concurrent
// callers share launches instead of serialising on the engine (the reference's counterpart is the OpenMP loop over blocks,
// src/smooth.cpp:1931, whose workers each own a private abpoa_t).
struct Coalescer {
    struct Batch {
        poa_b200_params_t params{};
        std::vector<int64_t> bso{0}, so{0};
        std::vector<int32_t> lens, wts;
        std::vector<uint8_t> bases;
        std::vector<uint64_t> tickets;
    };
    struct Done { std::shared_ptr<poa_b200_result> res; int64_t block = 0; int rc = 0; std::string err; };
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::unique_ptr<Batch> open;            // accumulating
    std::deque<std::unique_ptr<Batch>> ready;  // closed, waiting for the GPU
    std::unordered_map<uint64_t, Done> done;
    std::unordered_map<uint64_t, int> where;   // ticket -> 0 open / 1 ready or running (absent once done and collected)
    uint64_t next_ticket = 1;
    int64_t max_blocks = 8192;              // close the open batch at this many blocks
    int waiting_on_open = 0;                // threads blocked in wait() on a ticket of the open batch
    bool running = false, stop = false, started = false;
    std::thread worker;
};

struct poa_b200_engine {
    Coalescer co;
    int device = 0;
    int n_sm = 0;
    cudaStream_t stream = nullptr;
    poa_b200_engine_opts_t opts{};
    std::mutex mu;
    int occ[4] = {16, 6, 3, 1};  // resident CTAs per SM of poa_b200_block_kernel<1 / 2 / 4 / 8> (occupancy calculator, engine_create)
    size_t total_mem = 0;
    std::shared_ptr<PinnedPool> pinned = std::make_shared<PinnedPool>();
    DevicePool dev_pool;
};

struct poa_b200_result {
    int64_t n_blocks = 0;
    std::vector<int> hdr;                 // n_blocks * HDR_WORDS
    std::vector<int> arena_of;            // per block: which arena holds its body
    std::vector<int *> arenas;            // pinned host copies
    std::vector<size_t> arena_caps;       // bytes, for the pool (0 = not pooled)
    std::vector<int> owned;               // storage of a result built by poa_b200_result_from_parts
    std::shared_ptr<PinnedPool> pinned;
    std::vector<unsigned long long> arena_words;
    poa_b200_stats_t stats{};
    int emit_cigar = 0;
    // A result handed out by poa_b200_wait_block() is a window of ONE block onto the (shared) result of the batch its block
    // was coalesced into: `parent` owns the memory, block 0 of this object is block `parent_block` of the parent.
    std::shared_ptr<poa_b200_result> parent;
    int64_t parent_block = 0;
    // Block bodies are stored narrow and run-length coded (WireLayout, poa_core.cuh); a block is decoded into the view's flat
    // int32 arrays the first time it is looked at, lock-free (threads racing on one block both decode, one copy is kept).
    mutable std::unique_ptr<std::atomic<poa::DecodedBlock *>[]> decoded;
    void init_decoded() {
        decoded.reset(new std::atomic<poa::DecodedBlock *>[(size_t)std::max<int64_t>(n_blocks, 1)]);
        for (int64_t i = 0; i < std::max<int64_t>(n_blocks, 1); ++i) decoded[(size_t)i].store(nullptr, std::memory_order_relaxed);
    }
    ~poa_b200_result() {
        if (decoded) for (int64_t i = 0; i < std::max<int64_t>(n_blocks, 1); ++i) delete decoded[(size_t)i].load(std::memory_order_relaxed);
    }
};

struct poa_b200_graph {
    std::vector<int32_t> node_id, edge_from, edge_to, path_node;
    std::vector<char> node_base;
    std::vector<int64_t> path_off;
    std::vector<int64_t> seq_off;  // final graphs only: node k spells node_base[seq_off[k] .. seq_off[k+1])
};

struct poa_b200_batch {
    poa_b200_engine *eng = nullptr;
    poa_b200_params_t params{};
    DevParams dp{};
    int64_t n_blocks = 0, n_seqs = 0, n_bases = 0;
    // host copies of the index arrays (needed to size things and for retries)
    std::vector<long long> h_block_seq_off;
    std::vector<long long> h_block_bases;  // total bases per block
    std::vector<int> h_block_maxlen;
    std::vector<int> h_block_minlen;      // shortest sequence of the block
    std::vector<long long> h_block_excess;  // sum over sequences of max(0, length - median length): bases that long insertions add
    // device input
    long long *d_block_seq_off = nullptr, *d_seq_off = nullptr;
    int *d_seq_len = nullptr, *d_weight = nullptr, *d_order = nullptr;
    uint8_t *d_bases = nullptr;
    // device output / control
    int *d_hdr = nullptr, *d_counter = nullptr;
    unsigned long long *d_arena_used = nullptr, *d_phase = nullptr;
    std::vector<Arena> arenas;
    std::vector<int> arena_of;
    std::vector<int> h_hdr;
    std::vector<long long> need_words;  // per block: body words the kernel reported when the block overflowed an arena (0 = unknown)
    // workspace
    char *d_ws = nullptr;
    char *d_inputs = nullptr;  // one pooled allocation holding every d_* array above
    size_t inputs_cap = 0;
    long long ws_bytes = 0;
    WsLayout layout{};
    int n_ctas = 0, nw = 1;
    int n_pending = 0;  // blocks in the launch in flight
    bool launched = false, finished = false;
    std::string retry_error;  // why a re-run of overflowed blocks could not be launched (those blocks keep their error status)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    poa_b200_stats_t stats{};
};

namespace {

template <int NW>
cudaError_t launch_nw(int n_ctas, cudaStream_t st, const DevParams &P, const DevBatch &B, const WsLayout &L, char *ws, const DevOut &O) {
    // shared memory, not L1, is what limits resident POA blocks per SM (L1 hit rate of the row slab is ~5 %): ask for the
    // largest shared-memory carve-out; set per launch because the attribute is per device context
    cudaFuncSetAttribute(poa_b200_block_kernel<NW>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    poa_b200_block_kernel<NW><<<n_ctas, NW * 32, NW == 1 ? P16_SMEM_BYTES : p16_mw_smem_bytes<NW>(), st>>>(P, B, L, ws, O);
    return cudaGetLastError();
}

cudaError_t launch_kernel(int nw, int n_ctas, cudaStream_t st, const DevParams &P, const DevBatch &B, const WsLayout &L, char *ws, const DevOut &O) {
    switch (nw) {
        case 1: return launch_nw<1>(n_ctas, st, P, B, L, ws, O);
        case 2: return launch_nw<2>(n_ctas, st, P, B, L, ws, O);
        case 4: return launch_nw<4>(n_ctas, st, P, B, L, ws, O);
        default: return launch_nw<8>(n_ctas, st, P, B, L, ws, O);
    }
}

void free_batch_device(poa_b200_batch *b) {
    if (b->d_inputs) { b->eng->dev_pool.give(b->d_inputs, b->inputs_cap); b->d_inputs = nullptr; }  // all d_* input arrays live in it
    b->eng->dev_pool.give(b->d_ws, (size_t)b->ws_bytes); b->d_ws = nullptr;
    for (auto &a : b->arenas) b->eng->dev_pool.give(a.d, a.cap_bytes);
    b->arenas.clear();
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
}

struct Sizing {
    long long nmax, max_bases, max_len, max_seq, pool_growth, slab_bytes;
};

// Workspace sizing for a set of blocks.  level 0 = typical (fast path), 1 = 4x rows, 2 = 16x rows at full width,
// 3 (LEVEL_WORST) = worst case (one row per input base, full width).
constexpr int LEVEL_WORST = 3;
Sizing size_for(const poa_b200_batch *b, const std::vector<int> &blocks, int level, double rows_factor) {
    long long max_bases = 1, max_len = 1, max_seq = 1, max_excess = 0, max_spread = 0;
    for (int id : blocks) {
        max_bases = std::max(max_bases, b->h_block_bases[id]);
        max_len = std::max<long long>(max_len, b->h_block_maxlen[id]);
        max_seq = std::max<long long>(max_seq, b->h_block_seq_off[id + 1] - b->h_block_seq_off[id]);
        max_excess = std::max(max_excess, b->h_block_excess[id]);
        max_spread = std::max<long long>(max_spread, b->h_block_maxlen[id] - b->h_block_minlen[id]);
    }
    const long long worst_nodes = std::max<long long>(max_bases + 2, max_seq + 2);
    long long nmax, rows;
    const int wb = b->dp.local ? -1 : b->dp.wb;
    long long width = max_len + 1;
    if (level >= LEVEL_WORST) { nmax = worst_nodes; rows = worst_nodes; }
    else {
        // graph rows per query base: ~1.5 for 32 sequences at 2 % divergence, ~3.6 for 256: deep blocks get a
        // proportionally larger first guess so that they are not all re-run
        double f = rows_factor * (level == 1 ? 4.0 : (level == 2 ? 16.0 : 1.0)) * (1.0 + (double)std::max<long long>(0, max_seq - 32) / 128.0);
        // Long insertions / deletions (real blocks; the long-indel variant of the benchmark re-ran 23 % of its blocks before
        // this term existed): an insertion of d bases adds d rows and lengthens its sequence by d, so the bases in excess of
        // the block's median length estimate the extra rows, and the length spread the extra band width (the band follows
        // both the longest-path position and the arg-max columns, which an indel of d pulls d columns apart -- on the rows
        // near it only, hence a fraction of the spread).  Sized from that variant's measured cells per row; overflow still
        // falls back to the automatic re-run.
        rows = std::min<long long>(worst_nodes, (long long)(f * max_len) + 64 + max_excess);
        nmax = std::min<long long>(worst_nodes, rows + max_len / 2 + 64);
        if (wb >= 0) {
            long long w = wb + (long long)(b->dp.wf * max_len);
            width = level >= 2 ? max_len + 1 : std::min<long long>(max_len + 1, 2 * w + 1 + (level == 1 ? max_len / 2 : max_len / 8) + 32 + max_spread / 2);
        }
    }
    nmax = std::max<long long>(nmax, 1024);
    const long long vecs_per_row = width / 8 + 2;
    // int16 rows unless the block can reach the int32 regime (abpoa_align_simd.c:1293-1302)
    const long long len = std::max(max_len, nmax);
    // ... unless 16-bit cells provably hold every real value however long the graph gets (p16_safe_for_long_graph, poa_core.cuh:
    // monotone in the query length, so the block's longest sequence decides for all of its alignments)
    const bool may32 = std::max<long long>(max_len * b->dp.match, len * b->dp.e1 + b->dp.o1) > (long long)INT16_MAX - b->dp.min_mis - b->dp.oe1 - b->dp.oe2
                       && !(b->dp.p16_ok && p16_safe_for_long_graph(b->dp, max_len, max_len));
    Sizing s;
    s.nmax = nmax; s.max_bases = max_bases; s.max_len = max_len; s.max_seq = max_seq;
    const long long edges = std::min<long long>(max_bases + max_seq, level >= LEVEL_WORST ? (1LL << 60) : 3 * nmax);
    s.pool_growth = 8 * edges + 64;
    const int gen_planes = b->dp.gap_mode == 0 ? 3 : (b->dp.gap_mode == 1 ? 2 : 1);  // stored planes of the generic fill (H + E planes)
    long long row_bytes = vecs_per_row * gen_planes * (may32 ? 32 : 16);
    if (b->dp.p16_ok) {  // chunked rows of the packed 16-bit fill: whole 256-column chunks, P16_PLANES (H, E1, E2) x 512 B each
        const long long chunk_bytes = (long long)P16_PLANES * P16_CPB;
        long long p16_bytes = (width / 256 + 2) * chunk_bytes;  // worst case: a partial chunk at either end
        if (level == 0) {
            // first guess: a band-wide row touches band/256 + 1 chunks on average (measured 3.46 for 743-column bands,
            // whose rows are clipped at the matrix edges); 12 % headroom, overflow is retried at the next level
            const long long band = wb >= 0 ? std::min<long long>(max_len + 1, 2 * (wb + (long long)(b->dp.wf * max_len)) + 1 + max_spread / 3) : max_len + 1;
            p16_bytes = (long long)((band / 256.0 + 1.0) * 1.12 * (double)chunk_bytes);
        }
        // With match = 1 (smoothxg's default and every -a preset) p16_eligible() holds whenever the int16 test of
        // abpoa_align_simd.c:1293-1302 does, so a block that cannot reach the int32 regime is sized for packed rows alone;
        // the generic rows only matter for int32 rows.  A misjudged block is re-run at the next level.
        row_bytes = (level == 0 && !may32) ? p16_bytes : std::max(row_bytes, p16_bytes);
    }
    s.slab_bytes = rows * row_bytes;
    return s;
}

int run_launch(poa_b200_batch *b, const std::vector<int> &blocks, int level, cudaStream_t st, bool record_events) {
    poa_b200_engine *eng = b->eng;
    const double rows_factor = eng->opts.slab_rows_factor > 0 ? eng->opts.slab_rows_factor : 1.7;
    Sizing sz = size_for(b, blocks, level, rows_factor);
    WsLayout L;
    make_layout(L, sz.nmax, sz.max_bases, sz.max_len, sz.max_seq, sz.pool_growth, sz.slab_bytes, b->dp.emit_cigar);
    // How many warps per POA block, how many resident blocks per SM.  One warp per block (fill_p16) is by far the most efficient
    // use of the machine -- the per-row part of the work is serial and every further warp repeats it -- so several warps per
    // block only pay when the batch cannot fill the resident single-warp slots anyway: measured per-block time on 16 x 1 kb
    // blocks 62.7 ms with one warp, 66.6 with two, 47 with four (profiles/bench_r02_mw_*).  So: one warp while the blocks fill
    // the four-warp slots, else four, else eight (deep blocks: one block per SM, only row latency matters).  Resident CTAs per
    // SM come from the occupancy calculator of the instantiation actually launched (registers differ: 128 / 168 / 168 / 183).
    int nw = eng->opts.warps_per_block;
    if (nw != 1 && nw != 2 && nw != 4 && nw != 8) {
        const long long nb_ = (long long)blocks.size();
        nw = 1;
        if (b->dp.p16_ok) {
            if (nb_ <= (long long)eng->n_sm * eng->occ[3]) nw = 8;
            else if (nb_ <= (long long)eng->n_sm * eng->occ[2]) nw = 4;
        } else {
            // generic fill (int32 / affine / linear): its vectors are dealt to all threads of the block
            while (nw < 8 && (long long)nw * nb_ < 16LL * eng->n_sm) nw *= 2;
        }
    }
    const int occ = eng->occ[nw == 1 ? 0 : (nw == 2 ? 1 : (nw == 4 ? 2 : 3))];
    int per_sm = eng->opts.ctas_per_sm > 0 ? eng->opts.ctas_per_sm : std::max(1, occ);
    per_sm = std::min(per_sm, 32);
    long long n_ctas = std::min<long long>((long long)blocks.size(), (long long)eng->n_sm * per_sm);
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    long long budget = eng->opts.device_mem_budget > 0 ? eng->opts.device_mem_budget
                     : (long long)((double)(free_b + (size_t)b->ws_bytes + eng->dev_pool.pooled_bytes()) * 0.70);
    if (L.stride > budget) return set_err(POA_B200_ENOMEM, "one block's workspace exceeds the device memory budget");
    n_ctas = std::max<long long>(1, std::min<long long>(n_ctas, budget / std::max<long long>(L.stride, 1)));
    const long long need = n_ctas * L.stride;
    if (need > b->ws_bytes) {
        if (b->d_ws) { eng->dev_pool.give(b->d_ws, (size_t)b->ws_bytes); b->d_ws = nullptr; b->ws_bytes = 0; }
        size_t cap = 0;
        b->d_ws = (char *)eng->dev_pool.take((size_t)need, &cap);
        if (!b->d_ws) return set_err(POA_B200_ENOMEM, "workspace cudaMalloc failed");
        b->ws_bytes = (long long)cap;
    }
    b->layout = L; b->n_ctas = (int)n_ctas; b->nw = nw;
    // arena for this launch
    long long est_words = 0;
    for (int id : blocks) {
        long long tb = b->h_block_bases[id], ns = b->h_block_seq_off[id + 1] - b->h_block_seq_off[id], ml = b->h_block_maxlen[id];
        if (b->need_words[(size_t)id] > 0) { est_words += b->need_words[(size_t)id]; continue; }  // overflowed an arena before: the kernel reported its exact size
        long long n_est = level >= LEVEL_WORST ? tb + 2 : std::min<long long>(tb + 2, (level == 0 ? 3 : (level == 1 ? 8 : 24)) * ml + 64);
        // narrow, run-length coded bodies (WireLayout): ~4.5 words per node, a word per path run; level 0 assumes paths of few
        // runs, an overflow is re-run with the exact size the kernel reports
        long long wds = level == 0 ? 6 * n_est + tb / 4 + 5 * ns + 64 : 12 * n_est + 2 * tb + 5 * ns + 64;
        if (b->dp.out_msa) wds += (ns + 1) * n_est / 4 + 8;
        if (b->dp.emit_cigar) wds += 2 * (tb + ns * n_est);
        est_words += wds;
    }
    est_words = est_words + est_words / 4 + 1024;
    Arena ar;
    ar.cap = (unsigned long long)est_words;
    ar.d = (int *)eng->dev_pool.take((size_t)ar.cap * 4, &ar.cap_bytes);
    if (!ar.d) return set_err(POA_B200_ENOMEM, "arena cudaMalloc failed");
    b->arenas.push_back(ar);
    // order: most expensive first
    std::vector<int> order(blocks);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        double cx = (double)b->h_block_bases[x] * (double)b->h_block_bases[x];
        double cy = (double)b->h_block_bases[y] * (double)b->h_block_bases[y];
        return cx > cy;
    });
    CU(cudaMemcpyAsync(b->d_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(b->d_counter, 0, sizeof(int), st));
    CU(cudaMemsetAsync(b->d_arena_used, 0, sizeof(unsigned long long), st));
    DevBatch B;
    B.block_seq_off = b->d_block_seq_off; B.seq_len = b->d_seq_len; B.seq_off = b->d_seq_off;
    B.bases = b->d_bases; B.weight = b->d_weight; B.order = b->d_order; B.n_order = (int)order.size();
    DevOut O;
    O.hdr = b->d_hdr; O.arena = ar.d; O.arena_used = b->d_arena_used; O.arena_cap = ar.cap;
    O.phase = b->d_phase; O.counter = b->d_counter;
    CU(cudaEventRecord(b->ev0, st));  // every launch is timed: re-runs of overflowed blocks count towards stats.kernel_ms
    CU(launch_kernel(nw, (int)n_ctas, st, b->dp, B, L, b->d_ws, O));
    CU(cudaEventRecord(b->ev1, st));
    b->n_pending = (int)order.size();
    b->stats.kernel_launches += 1;
    if (record_events) { b->stats.n_ctas = (int)n_ctas; b->stats.warps_per_block = nw; b->stats.workspace_bytes = need; }
    return POA_B200_OK;
}

// after a launch has completed: fetch headers, note which arena each finished block used, list failures
int collect(poa_b200_batch *b, const std::vector<int> &blocks, cudaStream_t st, std::vector<int> &failed_ws, std::vector<int> &failed_arena) {
    CU(cudaMemcpyAsync(b->h_hdr.data(), b->d_hdr, b->h_hdr.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    unsigned long long used = 0;
    CU(cudaMemcpyAsync(&used, b->d_arena_used, sizeof(used), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    Arena &ar = b->arenas.back();
    ar.used = std::min(used, ar.cap);
    const int ai = (int)b->arenas.size() - 1;
    for (int id : blocks) {
        int st_ = b->h_hdr[(size_t)id * HDR_WORDS + H_STATUS];
        if (st_ == ST_OK) b->arena_of[id] = ai;
        else if (st_ == ST_ESLAB) failed_ws.push_back(id);
        else if (st_ == ST_EARENA) {
            failed_arena.push_back(id);
            const int *h = &b->h_hdr[(size_t)id * HDR_WORDS];
            b->need_words[(size_t)id] = (long long)((unsigned long long)(unsigned)h[H_OFF_LO] | ((unsigned long long)(unsigned)h[H_OFF_HI] << 32));
        }
    }
    return POA_B200_OK;
}

// one past the last arena word of a finished block's body, from its header; ~0 if the header is not a valid one
unsigned long long block_body_end(const int *h) {
    if (!wire_header_ok(h)) return ~0ull;
    const unsigned long long off = (unsigned long long)(unsigned)h[H_OFF_LO] | ((unsigned long long)(unsigned)h[H_OFF_HI] << 32);
    return off + (unsigned long long)(unsigned)h[H_BODY_WORDS];
}

int finish_locked(poa_b200_batch *b, cudaStream_t st) {
    if (!b->launched) return set_err(POA_B200_EARG, "batch was not launched");
    if (b->finished) return POA_B200_OK;
    std::vector<int> blocks((size_t)b->n_blocks);
    for (int64_t i = 0; i < b->n_blocks; ++i) blocks[(size_t)i] = (int)i;
    int level = 0;
    for (int round = 0; round < 8 && !blocks.empty(); ++round) {
        std::vector<int> f_ws, f_ar;
        int rc = collect(b, blocks, st, f_ws, f_ar);
        if (rc) return rc;
        {   // collect() has synchronised the stream: ev0 / ev1 bracket the launch just collected (main launch or a re-run)
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
            b->stats.kernel_ms = round == 0 ? ms : b->stats.kernel_ms + ms;
        }
        if (f_ws.empty() && f_ar.empty()) { blocks.clear(); break; }
        // workspace overflow: next sizing level.  Arena overflow alone keeps the workspace level: the re-run's arena is sized
        // from the exact body sizes the kernel wrote into the failed blocks' headers (need_words).
        if (!f_ws.empty()) { if (level == LEVEL_WORST) break; ++level; }
        std::vector<int> again(f_ws);
        again.insert(again.end(), f_ar.begin(), f_ar.end());
        std::sort(again.begin(), again.end());
        b->stats.retried_blocks += (int)again.size();
        rc = run_launch(b, again, level, st, false);
        while (rc == POA_B200_ENOMEM && !f_ws.empty() && level < LEVEL_WORST) rc = run_launch(b, again, ++level, st, false);  // a huge level-2 slab may not fit where the worst case does not either; try anyway
        blocks = again;
        if (rc) {
            // The re-run could not be launched (out of device memory, CUDA error).  Blocks that did finish keep their
            // results: the remaining ones stay marked ESLAB / EARENA and the call reports POA_B200_EBLOCK.
            b->retry_error = poa::g_last_error;
            blocks.clear();
            break;
        }
    }
    if (!blocks.empty()) {
        std::vector<int> f_ws, f_ar;
        int rc = collect(b, blocks, st, f_ws, f_ar);
        if (rc) return rc;
    }
    unsigned long long ph[PH_N];
    CU(cudaMemcpy(ph, b->d_phase, sizeof(ph), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 8; ++k) b->stats.phase_cycles[k] = (int64_t)ph[k];
    long long cells = 0, edges = 0;
    for (int64_t i = 0; i < b->n_blocks; ++i) {
        const int *h = &b->h_hdr[(size_t)i * HDR_WORDS];
        if (h[H_STATUS] != ST_OK) continue;
        cells += (long long)((unsigned long long)(unsigned)h[H_INBAND_LO] | ((unsigned long long)(unsigned)h[H_INBAND_HI] << 32));
        edges += (long long)((unsigned long long)(unsigned)h[H_EDGE_LO] | ((unsigned long long)(unsigned)h[H_EDGE_HI] << 32));
    }
    b->stats.inband_cells = cells; b->stats.edge_row_cells = edges;
    b->finished = true;
    return POA_B200_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int poa_b200_abi_version(void) { return POA_B200_ABI_VERSION; }

const char *poa_b200_strerror(int code) {
    switch (code) {
        case POA_B200_OK: return "ok";
        case POA_B200_ESLAB: return "DP workspace exhausted";
        case POA_B200_EARENA: return "result arena exhausted";
        case POA_B200_EINTERNAL: return "traceback dead end";
        case POA_B200_EUNSUP: return "unsupported parameters";
        case POA_B200_EBLOCK: return "one or more blocks failed";
        case POA_B200_ECUDA: return "CUDA error";
        case POA_B200_EARG: return "bad argument";
        case POA_B200_ENOMEM: return "out of memory";
        default: return "unknown";
    }
}

const char *poa_b200_last_error(void) { return g_last_error.c_str(); }

void poa_b200_encode_bases(const char *ascii, int64_t n, uint8_t *codes) {
    static const struct Table {
        uint8_t t[256];
        Table() {
            for (int i = 0; i < 256; ++i) t[i] = 4;
            t[0] = 0; t[1] = 1; t[2] = 2; t[3] = 3;
            t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2;
            t['T'] = t['t'] = t['U'] = t['u'] = 3;
        }
    } tab;
    const uint8_t *t = tab.t;  // local copy of the pointer: no guard check of the function-local static inside the loop
    int64_t i = 0;
    for (; i + 8 <= n; i += 8) {  // eight independent lookups per iteration, one 8-byte store
        uint64_t w;
        memcpy(&w, ascii + i, 8);
        const uint64_t o = (uint64_t)t[w & 0xff] | (uint64_t)t[(w >> 8) & 0xff] << 8 | (uint64_t)t[(w >> 16) & 0xff] << 16 | (uint64_t)t[(w >> 24) & 0xff] << 24
                         | (uint64_t)t[(w >> 32) & 0xff] << 32 | (uint64_t)t[(w >> 40) & 0xff] << 40 | (uint64_t)t[(w >> 48) & 0xff] << 48 | (uint64_t)t[w >> 56] << 56;
        memcpy(codes + i, &o, 8);
    }
    for (; i < n; ++i) codes[i] = t[(unsigned char)ascii[i]];
}

int poa_b200_engine_create(int device, const poa_b200_engine_opts_t *opts, poa_b200_engine_t **out) {
    if (!out) return set_err(POA_B200_EARG, "out is NULL");
    *out = nullptr;
    int n_dev = 0;
    CU(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) return set_err(POA_B200_EARG, "no such CUDA device");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return set_err(POA_B200_ECUDA, "this library holds sm_100a code only; device is not Blackwell");
    poa_b200_engine *e = new (std::nothrow) poa_b200_engine();
    if (!e) return set_err(POA_B200_ENOMEM, "engine alloc");
    e->device = device; e->n_sm = prop.multiProcessorCount; e->total_mem = prop.totalGlobalMem;
    if (opts) e->opts = *opts;
    CU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    {
        cudaFuncSetAttribute(poa_b200_block_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(poa_b200_block_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(poa_b200_block_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(poa_b200_block_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int o = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, poa_b200_block_kernel<1>, 32, P16_SMEM_BYTES) == cudaSuccess && o > 0) e->occ[0] = o;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, poa_b200_block_kernel<2>, 64, p16_mw_smem_bytes<2>()) == cudaSuccess && o > 0) e->occ[1] = o;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, poa_b200_block_kernel<4>, 128, p16_mw_smem_bytes<4>()) == cudaSuccess && o > 0) e->occ[2] = o;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, poa_b200_block_kernel<8>, 256, p16_mw_smem_bytes<8>()) == cudaSuccess && o > 0) e->occ[3] = o;
        cudaGetLastError();
    }
    *out = e;
    return POA_B200_OK;
}

int poa_b200_engine_trim(poa_b200_engine_t *eng) {
    if (!eng) return set_err(POA_B200_EARG, "NULL engine");
    std::lock_guard<std::mutex> lk(eng->mu);
    CU(cudaSetDevice(eng->device));
    eng->dev_pool.trim();
    eng->pinned->trim();
    return POA_B200_OK;
}

void poa_b200_engine_destroy(poa_b200_engine_t *eng) {
    if (!eng) return;
    {
        std::unique_lock<std::mutex> lk(eng->co.mu);
        eng->co.stop = true;
        eng->co.cv_work.notify_all();
    }
    if (eng->co.worker.joinable()) eng->co.worker.join();
    eng->co.done.clear();
    cudaSetDevice(eng->device);
    if (eng->stream) cudaStreamDestroy(eng->stream);
    delete eng;
}

int poa_b200_batch_upload(poa_b200_engine_t *eng, const poa_b200_params_t *params, int64_t n_blocks,
                          const int64_t *block_seq_off, const int32_t *seq_len, const int64_t *seq_off,
                          const uint8_t *bases, const int32_t *weight, poa_b200_batch_t **out) {
    if (!eng || !params || !out || n_blocks < 0 || (n_blocks > 0 && (!block_seq_off || !seq_off)))
        return set_err(POA_B200_EARG, "NULL argument");
    if (n_blocks > 0 && block_seq_off[n_blocks] > 0 && (!seq_len || !weight)) return set_err(POA_B200_EARG, "seq_len / weight is NULL");
    if (n_blocks > 0 && block_seq_off[n_blocks] > 0 && seq_off[block_seq_off[n_blocks]] > 0 && !bases) return set_err(POA_B200_EARG, "bases is NULL");
    *out = nullptr;
    int rc = check_params(*params);
    if (rc) return rc;
    if (n_blocks > INT32_MAX / HDR_WORDS) return set_err(POA_B200_EARG, "too many blocks in one batch");
    std::lock_guard<std::mutex> lk(eng->mu);
    CU(cudaSetDevice(eng->device));
    poa_b200_batch *b = new (std::nothrow) poa_b200_batch();
    if (!b) return set_err(POA_B200_ENOMEM, "batch alloc");
    b->eng = eng; b->params = *params; b->n_blocks = n_blocks;
    build_params(*params, eng->opts, b->dp);
    const int64_t n_seqs = n_blocks ? block_seq_off[n_blocks] : 0;
    const int64_t n_bases = n_seqs ? seq_off[n_seqs] : 0;
    b->n_seqs = n_seqs; b->n_bases = n_bases;
    b->h_block_seq_off.assign(block_seq_off, block_seq_off + (n_blocks ? n_blocks + 1 : 0));
    if (n_blocks == 0) b->h_block_seq_off.assign(1, 0);
    b->h_block_bases.resize((size_t)n_blocks); b->h_block_maxlen.resize((size_t)n_blocks);
    b->h_block_minlen.resize((size_t)n_blocks); b->h_block_excess.resize((size_t)n_blocks);
    std::vector<int> lens_tmp;
    for (int64_t i = 0; i < n_blocks; ++i) {
        int64_t s0 = block_seq_off[i], s1 = block_seq_off[i + 1];
        if (s1 < s0) { delete b; return set_err(POA_B200_EARG, "block_seq_off not monotone"); }
        int ml = 0, mn = s1 > s0 ? INT32_MAX : 0;
        for (int64_t s = s0; s < s1; ++s) {
            if (seq_len[s] < 0 || seq_off[s + 1] - seq_off[s] != seq_len[s]) { delete b; return set_err(POA_B200_EARG, "seq_off/seq_len mismatch"); }
            ml = std::max(ml, seq_len[s]); mn = std::min(mn, seq_len[s]);
        }
        long long excess = 0;
        if (s1 - s0 > 1) {
            lens_tmp.assign(seq_len + s0, seq_len + s1);
            std::nth_element(lens_tmp.begin(), lens_tmp.begin() + (long)(lens_tmp.size() / 2), lens_tmp.end());
            const int med = lens_tmp[lens_tmp.size() / 2];
            for (int64_t s = s0; s < s1; ++s) excess += std::max(0, seq_len[s] - med);
        }
        b->h_block_bases[(size_t)i] = seq_off[s1] - seq_off[s0];
        b->h_block_maxlen[(size_t)i] = ml; b->h_block_minlen[(size_t)i] = mn; b->h_block_excess[(size_t)i] = excess;
        if (b->h_block_bases[(size_t)i] > (1LL << 30)) { delete b; return set_err(POA_B200_EARG, "block too large"); }
    }
    auto fail = [&](int code) { free_batch_device(b); delete b; return code; };
#define CUB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err(POA_B200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); return fail(POA_B200_ECUDA); } } while (0)
    cudaStream_t st = eng->stream;
    CUB(cudaEventCreate(&b->ev0)); CUB(cudaEventCreate(&b->ev1));
    cudaEvent_t h0 = nullptr, h1 = nullptr;
    auto fail0 = fail;
    auto fail_ev = [&](int code) { if (h0) cudaEventDestroy(h0); if (h1) cudaEventDestroy(h1); return fail0(code); };
#undef CUB
#define CUB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err(POA_B200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); return fail_ev(POA_B200_ECUDA); } } while (0)
    CUB(cudaEventCreate(&h0)); CUB(cudaEventCreate(&h1));
    {
        // every per-batch input / bookkeeping array is carved from ONE pooled device allocation: cudaMalloc / cudaFree per
        // call cost up to hundreds of milliseconds each inside the end-to-end path (cudaFree synchronises the device)
        const size_t nb1 = (size_t)std::max<int64_t>(n_blocks, 1), ns1 = (size_t)std::max<int64_t>(n_seqs, 1);
        size_t off = 0;
        auto carve = [&off](size_t bytes) { const size_t at = off; off += (bytes + 255) & ~(size_t)255; return at; };
        const size_t o_bso = carve(sizeof(long long) * (size_t)(n_blocks + 1)), o_so = carve(sizeof(long long) * (size_t)(n_seqs + 1));
        const size_t o_sl = carve(sizeof(int) * ns1), o_wt = carve(sizeof(int) * ns1), o_ba = carve((size_t)std::max<int64_t>(n_bases, 1));
        const size_t o_or = carve(sizeof(int) * nb1), o_hd = carve(sizeof(int) * HDR_WORDS * nb1), o_ct = carve(sizeof(int));
        const size_t o_au = carve(sizeof(unsigned long long)), o_ph = carve(sizeof(unsigned long long) * PH_N);
        b->d_inputs = (char *)eng->dev_pool.take(off, &b->inputs_cap);
        if (!b->d_inputs) { set_err(POA_B200_ENOMEM, "input cudaMalloc failed"); return fail_ev(POA_B200_ENOMEM); }
        char *base = b->d_inputs;
        b->d_block_seq_off = (long long *)(base + o_bso); b->d_seq_off = (long long *)(base + o_so);
        b->d_seq_len = (int *)(base + o_sl); b->d_weight = (int *)(base + o_wt); b->d_bases = (uint8_t *)(base + o_ba);
        b->d_order = (int *)(base + o_or); b->d_hdr = (int *)(base + o_hd); b->d_counter = (int *)(base + o_ct);
        b->d_arena_used = (unsigned long long *)(base + o_au); b->d_phase = (unsigned long long *)(base + o_ph);
    }
    CUB(cudaMemsetAsync(b->d_phase, 0, sizeof(unsigned long long) * PH_N, st));
    CUB(cudaMemsetAsync(b->d_hdr, 0xff, sizeof(int) * HDR_WORDS * (size_t)std::max<int64_t>(n_blocks, 1), st));
    CUB(cudaEventRecord(h0, st));
    static const long long zero = 0;
    CUB(cudaMemcpyAsync(b->d_block_seq_off, n_blocks ? (const void *)block_seq_off : (const void *)&zero, sizeof(long long) * (size_t)(n_blocks + 1), cudaMemcpyHostToDevice, st));
    CUB(cudaMemcpyAsync(b->d_seq_off, n_seqs ? (const void *)seq_off : (const void *)&zero, sizeof(long long) * (size_t)(n_seqs + 1), cudaMemcpyHostToDevice, st));
    if (n_seqs) {
        CUB(cudaMemcpyAsync(b->d_seq_len, seq_len, sizeof(int) * (size_t)n_seqs, cudaMemcpyHostToDevice, st));
        CUB(cudaMemcpyAsync(b->d_weight, weight, sizeof(int) * (size_t)n_seqs, cudaMemcpyHostToDevice, st));
    }
    if (n_bases) {
        CUB(cudaMemcpyAsync(b->d_bases, bases, (size_t)n_bases, cudaMemcpyHostToDevice, st));
        poa_b200_sanitize_bases_kernel<<<std::max(1, eng->n_sm * 4), 256, 0, st>>>(b->d_bases, (long long)n_bases);
        CUB(cudaGetLastError());
    }
    CUB(cudaEventRecord(h1, st));
    CUB(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, h0, h1);
    cudaEventDestroy(h0); cudaEventDestroy(h1);
    b->stats.h2d_ms = ms;
    b->stats.h2d_bytes = (int64_t)(sizeof(long long) * (size_t)(n_blocks + 1 + n_seqs + 1) + 8 * (size_t)n_seqs + (size_t)n_bases);
    b->h_hdr.assign((size_t)n_blocks * HDR_WORDS, -1);
    b->arena_of.assign((size_t)n_blocks, -1);
    b->need_words.assign((size_t)n_blocks, 0);
#undef CUB
    *out = b;
    return POA_B200_OK;
}

int poa_b200_batch_launch(poa_b200_batch_t *b, void *stream) {
    if (!b) return set_err(POA_B200_EARG, "NULL batch");
    std::lock_guard<std::mutex> lk(b->eng->mu);
    CU(cudaSetDevice(b->eng->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : b->eng->stream;
    // a re-launch of the same batch recomputes everything (used by benchmarks)
    for (auto &a : b->arenas) b->eng->dev_pool.give(a.d, a.cap_bytes);
    b->arenas.clear();
    std::fill(b->arena_of.begin(), b->arena_of.end(), -1);
    b->stats.kernel_launches = 0; b->stats.retried_blocks = 0;
    CU(cudaMemsetAsync(b->d_phase, 0, sizeof(unsigned long long) * PH_N, st));
    b->finished = false;
    if (b->n_blocks == 0) { b->launched = true; b->finished = true; return POA_B200_OK; }
    std::vector<int> blocks((size_t)b->n_blocks);
    for (int64_t i = 0; i < b->n_blocks; ++i) blocks[(size_t)i] = (int)i;
    int rc = run_launch(b, blocks, 0, st, true);
    if (rc) return rc;
    b->launched = true;
    return POA_B200_OK;
}

int poa_b200_batch_finish(poa_b200_batch_t *b, void *stream) {
    if (!b) return set_err(POA_B200_EARG, "NULL batch");
    std::lock_guard<std::mutex> lk(b->eng->mu);
    CU(cudaSetDevice(b->eng->device));
    return finish_locked(b, stream ? (cudaStream_t)stream : b->eng->stream);
}

int poa_b200_batch_download(poa_b200_batch_t *b, void *stream, poa_b200_result_t **out) {
    if (!b || !out) return set_err(POA_B200_EARG, "NULL argument");
    *out = nullptr;
    std::lock_guard<std::mutex> lk(b->eng->mu);
    CU(cudaSetDevice(b->eng->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : b->eng->stream;
    int rc = finish_locked(b, st);
    if (rc) return rc;
    poa_b200_result *r = new (std::nothrow) poa_b200_result();
    if (!r) return set_err(POA_B200_ENOMEM, "result alloc");
    r->pinned = b->eng->pinned;
    r->n_blocks = b->n_blocks; r->hdr = b->h_hdr; r->arena_of = b->arena_of; r->emit_cigar = b->dp.emit_cigar;
    r->init_decoded();
    cudaEvent_t d0, d1;
    CU(cudaEventCreate(&d0)); CU(cudaEventCreate(&d1));
    CU(cudaEventRecord(d0, st));
    int64_t bytes = (int64_t)b->h_hdr.size() * 4;
    for (auto &a : b->arenas) {
        size_t nbytes = (size_t)std::max<unsigned long long>(a.used, 1) * 4, cap = 0;
        int *h = (int *)r->pinned->take(nbytes, &cap);
        if (!h) { poa_b200_result_free(r); return set_err(POA_B200_ENOMEM, "cudaMallocHost failed for the result buffer"); }
        r->arenas.push_back(h); r->arena_caps.push_back(cap); r->arena_words.push_back(a.used);
        if (a.used) CU(cudaMemcpyAsync(h, a.d, (size_t)a.used * 4, cudaMemcpyDeviceToHost, st));
        bytes += (int64_t)a.used * 4;
    }
    CU(cudaEventRecord(d1, st));
    CU(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, d0, d1);
    cudaEventDestroy(d0); cudaEventDestroy(d1);
    b->stats.d2h_ms = ms; b->stats.d2h_bytes = bytes;
    r->stats = b->stats;
    *out = r;
    for (int64_t i = 0; i < r->n_blocks; ++i)
        if (r->hdr[(size_t)i * HDR_WORDS + H_STATUS] != ST_OK)
            return set_err(POA_B200_EBLOCK, "one or more blocks failed; see per-block status" + (b->retry_error.empty() ? std::string() : " (re-run not launched: " + b->retry_error + ")"));
    return POA_B200_OK;
}

int poa_b200_batch_device_result(poa_b200_batch_t *b, int32_t arena_idx, const int32_t **d_hdr, const int32_t **d_arena,
                                 int64_t *arena_words, int32_t *n_arenas, const int32_t **block_arena) {
    if (!b || !b->finished) return set_err(POA_B200_EARG, "batch not finished");
    if (n_arenas) *n_arenas = (int32_t)b->arenas.size();
    if (d_hdr) *d_hdr = b->d_hdr;
    if (block_arena) *block_arena = b->arena_of.data();
    if (arena_idx < 0 || arena_idx >= (int)b->arenas.size()) {
        if (d_arena) *d_arena = nullptr;
        if (arena_words) *arena_words = 0;
        return b->arenas.empty() ? POA_B200_OK : set_err(POA_B200_EARG, "bad arena index");
    }
    if (d_arena) *d_arena = b->arenas[(size_t)arena_idx].d;
    if (arena_words) *arena_words = (int64_t)b->arenas[(size_t)arena_idx].used;
    return POA_B200_OK;
}

int poa_b200_result_from_parts(int64_t n_blocks, const int32_t *hdr, const int32_t *arena, int64_t arena_words, poa_b200_result_t **out) {
    if (!out || n_blocks < 0 || (n_blocks > 0 && !hdr) || arena_words < 0 || (arena_words > 0 && !arena)) return set_err(POA_B200_EARG, "bad argument");
    static_assert(POA_B200_HDR_WORDS == HDR_WORDS, "header size mismatch");
    poa_b200_result *r = new (std::nothrow) poa_b200_result();
    if (!r) return set_err(POA_B200_ENOMEM, "result alloc");
    r->n_blocks = n_blocks;
    r->init_decoded();
    r->hdr.assign(hdr, hdr + n_blocks * HDR_WORDS);
    r->arena_of.assign((size_t)n_blocks, 0);
    r->owned.assign(arena, arena + arena_words);
    r->arenas.push_back(r->owned.data()); r->arena_caps.push_back(0); r->arena_words.push_back((unsigned long long)arena_words);
    // validate offsets so a corrupt gather cannot make the accessors read out of bounds
    for (int64_t i = 0; i < n_blocks; ++i) {
        const int *h = &r->hdr[(size_t)i * HDR_WORDS];
        if (h[H_STATUS] != ST_OK) continue;
        if (block_body_end(h) > (unsigned long long)arena_words) { delete r; return set_err(POA_B200_EARG, "block body outside the arena"); }
    }
    *out = r;
    return POA_B200_OK;
}

int poa_b200_result_from_device_parts(poa_b200_engine_t *eng, int64_t n_blocks, const int32_t *hdr, const int32_t *d_arena,
                                       int64_t arena_words, void *stream, poa_b200_result_t **out) {
    if (!eng || !out || n_blocks < 0 || (n_blocks > 0 && !hdr) || arena_words < 0 || (arena_words > 0 && !d_arena)) return set_err(POA_B200_EARG, "bad argument");
    *out = nullptr;
    std::lock_guard<std::mutex> lk(eng->mu);
    CU(cudaSetDevice(eng->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : eng->stream;
    poa_b200_result *r = new (std::nothrow) poa_b200_result();
    if (!r) return set_err(POA_B200_ENOMEM, "result alloc");
    r->pinned = eng->pinned;
    r->n_blocks = n_blocks;
    r->init_decoded();
    r->hdr.assign(hdr, hdr + n_blocks * HDR_WORDS);
    r->arena_of.assign((size_t)n_blocks, 0);
    size_t cap = 0;
    int *h = (int *)r->pinned->take((size_t)std::max<int64_t>(arena_words, 1) * 4, &cap);
    if (!h) { delete r; return set_err(POA_B200_ENOMEM, "cudaMallocHost failed for the result buffer"); }
    r->arenas.push_back(h); r->arena_caps.push_back(cap); r->arena_words.push_back((unsigned long long)arena_words);
    cudaEvent_t d0, d1;
    CU(cudaEventCreate(&d0)); CU(cudaEventCreate(&d1));
    CU(cudaEventRecord(d0, st));
    if (arena_words) CU(cudaMemcpyAsync(h, d_arena, (size_t)arena_words * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(d1, st));
    CU(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, d0, d1);
    cudaEventDestroy(d0); cudaEventDestroy(d1);
    r->stats.d2h_ms = ms; r->stats.d2h_bytes = arena_words * 4 + n_blocks * HDR_WORDS * 4;
    for (int64_t i = 0; i < n_blocks; ++i) {  // same validation as poa_b200_result_from_parts
        const int *hh = &r->hdr[(size_t)i * HDR_WORDS];
        if (hh[H_STATUS] != ST_OK) continue;
        if (block_body_end(hh) > (unsigned long long)arena_words) { poa_b200_result_free(r); return set_err(POA_B200_EARG, "block body outside the arena"); }
    }
    *out = r;
    return POA_B200_OK;
}

void poa_b200_batch_free(poa_b200_batch_t *b) {
    if (!b) return;
    {
        std::lock_guard<std::mutex> lk(b->eng->mu);
        cudaSetDevice(b->eng->device);
        free_batch_device(b);
    }
    delete b;
}

int poa_b200_batch_stats(const poa_b200_batch_t *b, poa_b200_stats_t *s) {
    if (!b || !s) return set_err(POA_B200_EARG, "NULL argument");
    *s = b->stats;
    return POA_B200_OK;
}

int poa_b200_run_batch(poa_b200_engine_t *eng, const poa_b200_params_t *params, int64_t n_blocks,
                       const int64_t *block_seq_off, const int32_t *seq_len, const int64_t *seq_off,
                       const uint8_t *bases, const int32_t *weight, poa_b200_result_t **result) {
    if (!result) return set_err(POA_B200_EARG, "NULL result");
    *result = nullptr;
    poa_b200_batch_t *b = nullptr;
    static const bool trace = getenv("POA_B200_TRACE") != nullptr;  // host wall time of the four stages, to stderr
    const auto t0 = std::chrono::steady_clock::now();
    int rc = poa_b200_batch_upload(eng, params, n_blocks, block_seq_off, seq_len, seq_off, bases, weight, &b);
    if (rc) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    rc = poa_b200_batch_launch(b, nullptr);
    const auto t2 = std::chrono::steady_clock::now();
    if (rc == POA_B200_OK) rc = poa_b200_batch_download(b, nullptr, result);
    const auto t3 = std::chrono::steady_clock::now();
    poa_b200_batch_free(b);
    if (trace) {
        const auto t4 = std::chrono::steady_clock::now();
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point z) { return std::chrono::duration<double, std::milli>(z - a).count(); };
        fprintf(stderr, "[poa_b200] run_batch %lld blocks: upload %.2f ms, launch+wait %.2f ms, download %.2f ms, free %.2f ms\n",
                (long long)n_blocks, ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4));
    }
    return rc;
}

namespace {
void coalescer_main(poa_b200_engine *eng) {
    Coalescer &co = eng->co;
    std::unique_lock<std::mutex> lk(co.mu);
    for (;;) {
        // close the open batch when it is full, or when somebody is waiting for one of its blocks and the GPU would idle
        co.cv_work.wait(lk, [&] { return co.stop || !co.ready.empty() || (co.open && !co.open->tickets.empty() && co.waiting_on_open > 0); });
        if (co.stop && co.ready.empty() && !(co.open && !co.open->tickets.empty())) break;
        if (co.ready.empty() && co.open && !co.open->tickets.empty()) {
            for (uint64_t t : co.open->tickets) co.where[t] = 1;
            co.ready.push_back(std::move(co.open));
            co.waiting_on_open = 0;
        }
        if (co.ready.empty()) continue;
        std::unique_ptr<Coalescer::Batch> b = std::move(co.ready.front());
        co.ready.pop_front();
        co.running = true;
        lk.unlock();
        poa_b200_result_t *res = nullptr;
        const int64_t nb = (int64_t)b->tickets.size();
        const int rc = poa_b200_run_batch(eng, &b->params, nb, b->bso.data(), b->lens.data(), b->so.data(),
                                          b->bases.empty() ? nullptr : b->bases.data(), b->wts.data(), &res);
        const std::string err = poa::g_last_error;
        std::shared_ptr<poa_b200_result> holder;
        if (res) holder.reset(res, [](poa_b200_result *r) { poa_b200_result_free(r); });
        lk.lock();
        for (int64_t i = 0; i < nb; ++i) {
            Coalescer::Done d;
            d.res = holder; d.block = i;
            if (!res) { d.rc = rc ? rc : POA_B200_EINTERNAL; d.err = err; }
            else { const int st = res->hdr[(size_t)i * HDR_WORDS + H_STATUS]; d.rc = st == ST_OK ? POA_B200_OK : POA_B200_EBLOCK; if (st != ST_OK) d.err = err; }
            co.done.emplace(b->tickets[(size_t)i], std::move(d));
            co.where.erase(b->tickets[(size_t)i]);
        }
        co.running = false;
        co.cv_done.notify_all();
    }
}
}  // namespace

int poa_b200_submit_block(poa_b200_engine_t *eng, const poa_b200_params_t *params, int32_t n_seq,
                          const uint8_t *const *seqs, const int32_t *seq_lens, const int32_t *weights, uint64_t *ticket) {
    if (!eng || !params || !ticket || n_seq < 0 || (n_seq > 0 && (!seqs || !seq_lens || !weights))) return set_err(POA_B200_EARG, "bad block");
    int rc = check_params(*params);
    if (rc) return rc;
    for (int i = 0; i < n_seq; ++i) if (seq_lens[i] < 0 || (seq_lens[i] > 0 && !seqs[i])) return set_err(POA_B200_EARG, "bad sequence");
    Coalescer &co = eng->co;
    std::unique_lock<std::mutex> lk(co.mu);
    if (co.stop) return set_err(POA_B200_EARG, "engine is shutting down");
    if (!co.started) { co.worker = std::thread(coalescer_main, eng); co.started = true; }
    // one batch = one parameter set: a block with other parameters closes the open batch first
    if (co.open && !co.open->tickets.empty() && memcmp(&co.open->params, params, sizeof(*params)) != 0) {
        for (uint64_t t : co.open->tickets) co.where[t] = 1;
        co.ready.push_back(std::move(co.open));
        co.waiting_on_open = 0;
        co.cv_work.notify_all();
    }
    if (!co.open) { co.open.reset(new Coalescer::Batch()); }
    Coalescer::Batch &b = *co.open;
    if (b.tickets.empty()) b.params = *params;
    for (int i = 0; i < n_seq; ++i) {
        b.lens.push_back(seq_lens[i]); b.wts.push_back(weights[i]);
        if (seq_lens[i]) b.bases.insert(b.bases.end(), seqs[i], seqs[i] + seq_lens[i]);
        b.so.push_back((int64_t)b.bases.size());
    }
    b.bso.push_back((int64_t)b.lens.size());
    const uint64_t t = co.next_ticket++;
    b.tickets.push_back(t);
    co.where[t] = 0;
    *ticket = t;
    if ((int64_t)b.tickets.size() >= co.max_blocks) {
        for (uint64_t x : b.tickets) co.where[x] = 1;
        co.ready.push_back(std::move(co.open));
        co.waiting_on_open = 0;
        co.cv_work.notify_all();
    }
    return POA_B200_OK;
}

int poa_b200_wait_block(poa_b200_engine_t *eng, uint64_t ticket, poa_b200_result_t **result) {
    if (!eng || !result) return set_err(POA_B200_EARG, "NULL argument");
    *result = nullptr;
    Coalescer &co = eng->co;
    std::unique_lock<std::mutex> lk(co.mu);
    auto it = co.done.find(ticket);
    if (it == co.done.end()) {
        auto w = co.where.find(ticket);
        if (w == co.where.end()) return set_err(POA_B200_EARG, "unknown or already collected ticket");
        bool counted = false;
        while ((it = co.done.find(ticket)) == co.done.end()) {
            w = co.where.find(ticket);
            if (w != co.where.end() && w->second == 0 && !counted) { ++co.waiting_on_open; counted = true; co.cv_work.notify_all(); }
            co.cv_done.wait(lk);
        }
    }
    Coalescer::Done d = std::move(it->second);
    co.done.erase(it);
    lk.unlock();
    if (!d.res) return set_err(d.rc ? d.rc : POA_B200_EINTERNAL, d.err);
    poa_b200_result *r = new (std::nothrow) poa_b200_result();
    if (!r) return set_err(POA_B200_ENOMEM, "result alloc");
    r->n_blocks = 1; r->parent = d.res; r->parent_block = d.block; r->emit_cigar = d.res->emit_cigar;
    *result = r;
    return d.rc == POA_B200_OK ? POA_B200_OK : set_err(d.rc, d.err.empty() ? "block failed; see its status" : d.err);
}

int poa_b200_poa_block(poa_b200_engine_t *eng, const poa_b200_params_t *params, int32_t n_seq,
                       const uint8_t *const *seqs, const int32_t *seq_lens, const int32_t *weights,
                       poa_b200_result_t **result) {
    if (!result) return set_err(POA_B200_EARG, "NULL result");
    *result = nullptr;
    uint64_t t = 0;
    int rc = poa_b200_submit_block(eng, params, n_seq, seqs, seq_lens, weights, &t);
    if (rc) return rc;
    return poa_b200_wait_block(eng, t, result);
}

int poa_b200_block_graph(const poa_b200_block_view_t *v, int32_t padding_len, int32_t include_consensus, poa_b200_graph_t **out) {
    if (!v || !out || padding_len < 0) return set_err(POA_B200_EARG, "bad argument");
    *out = nullptr;
    if (v->status != POA_B200_OK) return set_err(POA_B200_EBLOCK, "block has no result");
    poa_b200_graph *g = new (std::nothrow) poa_b200_graph();
    if (!g) return set_err(POA_B200_ENOMEM, "graph alloc");
    const int n = v->n_node, ns = v->n_seq;
    g->path_off.assign(1, 0);
    if (n <= 2) {  // src/smooth.cpp:2451: nothing is built
        for (int i = 0; i < ns + (include_consensus ? 1 : 0); ++i) g->path_off.push_back(0);
        *out = g;
        return POA_B200_OK;
    }
    std::vector<int64_t> in_off((size_t)n + 1, 0), out_off((size_t)n + 1, 0);
    for (int i = 0; i < n; ++i) { in_off[(size_t)i + 1] = in_off[(size_t)i] + v->in_n[i]; out_off[(size_t)i + 1] = out_off[(size_t)i] + v->out_n[i]; }
    // read paths with the padding trimmed (:2520-2531), node coverage
    std::vector<char> covered((size_t)n, 0);
    {
        int64_t keep = 0, at = 0;
        for (int i = 0; i < ns; ++i) keep += std::max(0, v->path_len[i] - 2 * padding_len);
        g->path_node.reserve((size_t)keep + (size_t)std::max(v->cons_len, 0));
        g->path_off.reserve((size_t)ns + 2);
        for (int i = 0; i < ns; ++i) {
            const int len = v->path_len[i];
            for (int j = padding_len; j < len - padding_len; ++j) {
                const int id = v->path_node[at + j];
                g->path_node.push_back(id - 1);
                covered[(size_t)id] = 1;
            }
            g->path_off.push_back((int64_t)g->path_node.size());
            at += len;
        }
    }
    if (include_consensus) {  // :2534-2549: only nodes some read still covers
        for (int i = 0; i < v->cons_len; ++i) { const int id = v->cons_node[i]; if (covered[(size_t)id]) g->path_node.push_back(id - 1); }
        g->path_off.push_back((int64_t)g->path_node.size());
    }
    // Which edges survive: build_odgi_abPOA asks odgi for the edges of path depth < 1 and destroys them (:2559-2565), but
    // find_edges_exceeding_depth_limits (deps/odgi/src/algorithms/depth.cpp:17-51) only ever looks at edges some path step
    // walks, whose depth is >= 1 by construction, so with min_depth = 1 it returns nothing: NO edge is removed there.  Edges
    // disappear only together with an uncovered node (:2567-2573, destroy_handle).  So every POA edge between two covered
    // nodes is kept, walked or not -- and the unwalked ones matter: they keep unchop from merging their end nodes.
    // Kahn walk from the source in out_id order = node / edge creation order of build_odgi_abPOA (:2463-2511)
    static const char code2base[6] = {'A', 'C', 'G', 'T', 'N', '-'};
    std::vector<int> indeg((size_t)n), queue; queue.reserve((size_t)n);
    for (int i = 0; i < n; ++i) indeg[(size_t)i] = v->in_n[i];
    queue.push_back(0);
    for (size_t qh = 0; qh < queue.size(); ++qh) {
        const int cur = queue[qh];
        if (cur == 1) break;
        if (cur != 0) {
            if (covered[(size_t)cur]) {
                g->node_id.push_back(cur - 1);
                const int b = v->base[cur];
                g->node_base.push_back(code2base[b >= 0 && b < 5 ? b : 4]);
            }
            for (int64_t k = in_off[(size_t)cur]; k < in_off[(size_t)cur + 1]; ++k) {
                const int pre = v->in_id[k];
                if (pre == 0 || !covered[(size_t)pre] || !covered[(size_t)cur]) continue;
                g->edge_from.push_back(pre - 1); g->edge_to.push_back(cur - 1);
            }
        }
        for (int64_t k = out_off[(size_t)cur]; k < out_off[(size_t)cur + 1]; ++k) {
            const int o = v->out_id[k];
            if (--indeg[(size_t)o] == 0) queue.push_back(o);
        }
    }
    *out = g;
    return POA_B200_OK;
}

int poa_b200_block_final_graph(const poa_b200_block_view_t *v, int32_t padding_len, int32_t include_consensus, poa_b200_graph_t **out) {
    if (!out) return set_err(POA_B200_EARG, "bad argument");
    *out = nullptr;
    poa_b200_graph_t *g1 = nullptr;
    int rc = poa_b200_block_graph(v, padding_len, include_consensus, &g1);  // 1-bp nodes, path-walked edges, trimmed paths
    if (rc) return rc;
    std::unique_ptr<poa_b200_graph> one(g1);
    poa_b200_graph *g = new (std::nothrow) poa_b200_graph();
    if (!g) return set_err(POA_B200_ENOMEM, "graph alloc");
    const size_t n_path = one->path_off.size() - 1;
    g->path_off.assign(1, 0);
    g->seq_off.assign(1, 0);
    const int n1 = (int)one->node_id.size();
    if (n1 == 0) {
        for (size_t p = 0; p < n_path; ++p) g->path_off.push_back(0);
        *out = g;
        return POA_B200_OK;
    }
    // dense index of the kept 1-bp nodes (odgi id -> position in creation order)
    int max_id = 0;
    for (int x : one->node_id) max_id = std::max(max_id, x);
    std::vector<int> dense((size_t)max_id + 1, -1);
    for (int k = 0; k < n1; ++k) dense[(size_t)one->node_id[(size_t)k]] = k;
    // every path step as a dense node index, once (the three passes below walk plain arrays)
    const size_t n_steps = one->path_node.size();
    std::vector<int> dstep(n_steps);
    { const int32_t *pn = one->path_node.data(); const int *dn = dense.data(); int *ds = dstep.data(); for (size_t k = 0; k < n_steps; ++k) ds[k] = dn[pn[k]]; }
    const int64_t *poff = one->path_off.data();
    const int *ds = dstep.data();
    // unchop (deps/odgi/src/algorithms/unchop.cpp, simple_components.cpp:27-60, perfect_neighbors.cpp:10-100): two nodes merge when
    // every path step on the left one continues to the right one and every step on the right one comes from the left one
    // -- no path starts, ends or branches between them.  NONE = no step seen yet, MANY = several targets or a path end.
    const int NONE = -1, MANY = -2;
    std::vector<int> nxt((size_t)n1, NONE), prv((size_t)n1, NONE);
    for (size_t p = 0; p < n_path; ++p) {
        const int64_t a = poff[p], b = poff[p + 1];
        if (a == b) continue;
        prv[(size_t)ds[a]] = MANY;      // a path starts here
        nxt[(size_t)ds[b - 1]] = MANY;  // a path ends here
        int *nx = nxt.data(), *pv = prv.data();
        for (int64_t k = a; k + 1 < b; ++k) {
            const int u = ds[k], w = ds[k + 1];
            int &x = nx[u]; x = (x == NONE || x == w) ? w : MANY;
            int &y = pv[w]; y = (y == NONE || y == u) ? u : MANY;
        }
    }
    // ... and the two are joined by the only edge leaving the left and the only edge entering the right one (simple_components.cpp:
    // get_degree == 1 on both sides), counted over ALL kept edges, walked by a path or not
    std::vector<int> outdeg((size_t)n1, 0), indeg1((size_t)n1, 0);
    for (size_t e = 0; e < one->edge_from.size(); ++e) { ++outdeg[(size_t)dense[(size_t)one->edge_from[e]]]; ++indeg1[(size_t)dense[(size_t)one->edge_to[e]]]; }
    // link[u] = the node u is merged with on its right, or -1
    std::vector<int> link((size_t)n1, -1);
    std::vector<char> is_head((size_t)n1, 1);
    for (int u = 0; u < n1; ++u) {
        const int w = nxt[(size_t)u];
        if (w >= 0 && prv[(size_t)w] == u && outdeg[(size_t)u] == 1 && indeg1[(size_t)w] == 1) { link[(size_t)u] = w; is_head[(size_t)w] = 0; }
    }
    // merged nodes: chains of linked 1-bp nodes, discovered in creation order of their heads
    std::vector<int> comp((size_t)n1, -1);
    std::vector<int> head;  // first 1-bp node of every merged node
    for (int k = 0; k < n1; ++k) {
        if (!is_head[(size_t)k]) continue;
        const int c = (int)head.size();
        head.push_back(k);
        for (int u = k; u >= 0; u = link[(size_t)u]) comp[(size_t)u] = c;
    }
    const int nc = (int)head.size();
    // edges between merged nodes: every consecutive pair of path steps that crosses a chain boundary (src/smooth.cpp:590-606 creates
    // an edge for every such pair, consensus steps included); out-degrees are tiny, so duplicates are found by a short scan
    std::vector<std::vector<int>> outs((size_t)nc);
    size_t n_edges = 0;
    {
        const int *lk = link.data(), *cp = comp.data();
        for (size_t p = 0; p < n_path; ++p) {
            const int64_t a = poff[p], b = poff[p + 1];
            for (int64_t k = a; k + 1 < b; ++k) {
                const int u = ds[k], w = ds[k + 1];
                if (lk[u] == w) continue;
                std::vector<int> &o = outs[(size_t)cp[u]];
                const int cw = cp[w];
                if (std::find(o.begin(), o.end(), cw) == o.end()) { o.push_back(cw); ++n_edges; }
            }
        }
    }
    std::vector<std::pair<int, int>> edges;
    edges.reserve(n_edges);
    for (int c = 0; c < nc; ++c) { std::sort(outs[(size_t)c].begin(), outs[(size_t)c].end()); for (int w : outs[(size_t)c]) edges.emplace_back(c, w); }
    // topological order (src/smooth.cpp:557: apply_ordering(topological_order(...), compact ids)): Kahn's algorithm, ready nodes
    // taken in discovery order.  odgi's own tie-breaking keys on handle ranks that come out of a hash map after unchop, so
    // the reference's ids are not a function of the block; any topological order gives an isomorphic graph.
    std::vector<int> indeg((size_t)nc, 0), eoff((size_t)nc + 1, 0), rank((size_t)nc, -1), order;
    for (auto &e : edges) { ++indeg[(size_t)e.second]; ++eoff[(size_t)e.first + 1]; }
    for (int c = 0; c < nc; ++c) eoff[(size_t)c + 1] += eoff[(size_t)c];
    order.reserve((size_t)nc);
    for (int c = 0; c < nc; ++c) if (indeg[(size_t)c] == 0) order.push_back(c);
    for (size_t qh = 0; qh < order.size(); ++qh) {
        const int c = order[qh];
        for (int e = eoff[(size_t)c]; e < eoff[(size_t)c + 1]; ++e) if (--indeg[(size_t)edges[(size_t)e].second] == 0) order.push_back(edges[(size_t)e].second);
    }
    if ((int)order.size() != nc) { delete g; return set_err(POA_B200_EINTERNAL, "block graph is not acyclic"); }
    for (int r = 0; r < nc; ++r) rank[(size_t)order[(size_t)r]] = r;
    // nodes in rank order: id = rank + 1, sequence = the chain's bases
    g->node_id.resize((size_t)nc);
    for (int r = 0; r < nc; ++r) {
        g->node_id[(size_t)r] = r + 1;
        for (int u = head[(size_t)order[(size_t)r]]; u >= 0; u = link[(size_t)u]) g->node_base.push_back(one->node_base[(size_t)u]);
        g->seq_off.push_back((int64_t)g->node_base.size());
    }
    for (auto &e : edges) { g->edge_from.push_back(rank[(size_t)e.first] + 1); g->edge_to.push_back(rank[(size_t)e.second] + 1); }
    // paths: one step per merged node (a path that enters a chain walks all of it)
    {
        const char *ih = is_head.data(); const int *cp = comp.data(), *rk = rank.data();
        g->path_node.reserve(n_steps / 2 + 16);
        for (size_t p = 0; p < n_path; ++p) {
            const int64_t a = poff[p], b = poff[p + 1];
            for (int64_t k = a; k < b; ++k) { const int u = ds[k]; if (ih[u]) g->path_node.push_back(rk[cp[u]] + 1); }
            g->path_off.push_back((int64_t)g->path_node.size());
        }
    }
    *out = g;
    return POA_B200_OK;
}

int poa_b200_final_graph_view(const poa_b200_graph_t *g, poa_b200_final_graph_view_t *v) {
    if (!g || !v || g->seq_off.empty()) return set_err(POA_B200_EARG, "not a final graph");
    v->n_node = (int32_t)g->node_id.size(); v->seq_off = g->seq_off.data(); v->seq = g->node_base.data();
    v->n_edge = (int32_t)g->edge_from.size(); v->edge_from = g->edge_from.data(); v->edge_to = g->edge_to.data();
    v->n_path = (int32_t)g->path_off.size() - 1; v->path_off = g->path_off.data(); v->path_node = g->path_node.data();
    return POA_B200_OK;
}

int poa_b200_graph_view(const poa_b200_graph_t *g, poa_b200_graph_view_t *v) {
    if (!g || !v) return set_err(POA_B200_EARG, "NULL argument");
    v->n_node = (int32_t)g->node_id.size(); v->node_id = g->node_id.data(); v->node_base = g->node_base.data();
    v->n_edge = (int32_t)g->edge_from.size(); v->edge_from = g->edge_from.data(); v->edge_to = g->edge_to.data();
    v->n_path = (int32_t)g->path_off.size() - 1; v->path_off = g->path_off.data(); v->path_node = g->path_node.data();
    return POA_B200_OK;
}

void poa_b200_graph_free(poa_b200_graph_t *g) { delete g; }

int64_t poa_b200_result_n_blocks(const poa_b200_result_t *res) { return res ? res->n_blocks : 0; }

int poa_b200_result_block(const poa_b200_result_t *res, int64_t blk, poa_b200_block_view_t *v) {
    if (!res || !v || blk < 0 || blk >= res->n_blocks) return set_err(POA_B200_EARG, "bad block index");
    if (res->parent) return poa_b200_result_block(res->parent.get(), res->parent_block + blk, v);
    memset(v, 0, sizeof(*v));
    const int *h = &res->hdr[(size_t)blk * HDR_WORDS];
    v->status = h[H_STATUS];
    v->n_seq = h[H_N_SEQ];
    if (v->status != ST_OK) { v->cons_len = -1; v->msa_len = -1; return POA_B200_OK; }
    const int ai = res->arena_of[(size_t)blk];
    if (ai < 0 || ai >= (int)res->arenas.size()) return set_err(POA_B200_EARG, "block has no result body");
    const unsigned long long off = (unsigned long long)(unsigned)h[H_OFF_LO] | ((unsigned long long)(unsigned)h[H_OFF_HI] << 32);
    if (block_body_end(h) > res->arena_words[(size_t)ai]) return set_err(POA_B200_EARG, "block body outside the arena");
    std::atomic<poa::DecodedBlock *> &slot = res->decoded[(size_t)blk];
    poa::DecodedBlock *d = slot.load(std::memory_order_acquire);
    if (!d) {
        d = new (std::nothrow) poa::DecodedBlock();
        if (!d) return set_err(POA_B200_ENOMEM, "decode buffer");
        if (!wire_decode(h, res->arenas[(size_t)ai] + off, *d)) { delete d; return set_err(POA_B200_EARG, "corrupt block body"); }
        poa::DecodedBlock *expect = nullptr;
        if (!slot.compare_exchange_strong(expect, d, std::memory_order_acq_rel)) { delete d; d = expect; }
    }
    const int32_t *o = d->buf.data();
    const int n = h[H_N_NODE];
    const long long in_tot = h[H_IN_TOT], out_tot = h[H_OUT_TOT], aln_tot = h[H_ALN_TOT], path_tot = h[H_PATH_TOT], cig_tot = h[H_CIG_TOT];
    v->n_node = n; v->cons_len = h[H_CONS_LEN]; v->msa_len = h[H_MSA_LEN]; v->msa_rows = h[H_MSA_ROWS];
    v->base = o + d->base; v->in_n = o + d->in_n; v->in_id = o + d->in_id; v->in_w = o + d->in_w;
    v->out_n = o + d->out_n; v->out_id = o + d->out_id; v->out_w = o + d->out_w;
    v->aln_n = o + d->aln_n; v->aln_id = o + d->aln_id;
    v->path_len = o + d->path_len; v->path_node = o + d->path_node; v->cons_node = o + d->cons_node;
    v->best_score = o + d->best; v->n_cigar = o + d->ncig;
    v->msa = d->msa;
    v->cigar = reinterpret_cast<const uint64_t *>(d->cig);  // lo word first: little-endian uint64, 4-byte aligned
    v->in_total = in_tot; v->out_total = out_tot; v->aln_total = aln_tot; v->path_total = path_tot; v->cigar_total = cig_tot;
    v->inband_cells = (int64_t)((unsigned long long)(unsigned)h[H_INBAND_LO] | ((unsigned long long)(unsigned)h[H_INBAND_HI] << 32));
    return POA_B200_OK;
}

void poa_b200_result_release_block(const poa_b200_result_t *res, int64_t blk) {
    if (res && res->parent && blk >= 0 && blk < res->n_blocks) { poa_b200_result_release_block(res->parent.get(), res->parent_block + blk); return; }
    if (!res || blk < 0 || blk >= res->n_blocks || !res->decoded) return;
    delete res->decoded[(size_t)blk].exchange(nullptr, std::memory_order_acq_rel);
}

int poa_b200_result_block_hash(const poa_b200_result_t *res, int64_t blk, uint64_t *hash) {
    if (!hash) return set_err(POA_B200_EARG, "NULL argument");
    poa_b200_block_view_t v;
    int rc = poa_b200_result_block(res, blk, &v);
    if (rc) return rc;
    if (v.status != POA_B200_OK) return set_err(POA_B200_EBLOCK, "block has no result");
    // FNV-1a over what abPOA holds for the graph: node_n (int), then per node its base (one byte), out_id[] and
    // out_edge_weight[] (ints) -- the byte stream oracle/ref_shim.c:ref_poa_batch_timed hashes from abpoa_graph_t
    uint64_t h = 1469598103934665603ULL;
    auto mix = [&h](const void *d, size_t n) { const uint8_t *p = (const uint8_t *)d; for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ULL; } };
    const int32_t n = v.n_node;
    mix(&n, sizeof(n));
    int64_t at = 0;
    for (int i = 0; i < n; ++i) {
        const uint8_t b = (uint8_t)v.base[i];
        mix(&b, 1);
        mix(v.out_id + at, sizeof(int32_t) * (size_t)v.out_n[i]);
        mix(v.out_w + at, sizeof(int32_t) * (size_t)v.out_n[i]);
        at += v.out_n[i];
    }
    *hash = h;
    return POA_B200_OK;
}

int poa_b200_result_stats(const poa_b200_result_t *res, poa_b200_stats_t *s) {
    if (!res || !s) return set_err(POA_B200_EARG, "NULL argument");
    *s = res->parent ? res->parent->stats : res->stats;
    return POA_B200_OK;
}

void poa_b200_result_free(poa_b200_result_t *res) {
    if (!res) return;
    if (res->parent) { poa_b200_result_release_block(res->parent.get(), res->parent_block); delete res; return; }  // the last window frees the batch result
    for (size_t i = 0; i < res->arenas.size(); ++i) if (res->arena_caps[i]) res->pinned->give(res->arenas[i], res->arena_caps[i]);
    delete res;
}

}  // extern "C"
