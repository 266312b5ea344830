// 16 host threads call the per-block entry points of include/poa_b200.h concurrently, the way the OpenMP workers of smoothxg's
// block loop (reference src/smooth.cpp:1931) would: each thread keeps a window of outstanding tickets (submit, later wait).
// Prints the throughput of the batched call and of the coalesced per-block path and a checksum of both results.
//   usage: coalesce_test <n_blocks> <n_seqs> <len> <threads> <window>      (window 1 = synchronous poa_b200_poa_block)
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include "poa_b200.h"

static uint64_t rng_state = 88172645463325252ULL;
static inline uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 32); }

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: coalesce_test n_blocks n_seqs len threads window\n"); return 2; }
    const int nb = atoi(argv[1]), ns = atoi(argv[2]), L = atoi(argv[3]), nt = atoi(argv[4]), window = atoi(argv[5]);
    // synthetic blocks: a random base sequence, each copy mutated at 2 % (substitutions and 1-bp indels)
    std::vector<std::vector<std::vector<uint8_t>>> blocks((size_t)nb);
    for (auto &b : blocks) {
        std::vector<uint8_t> base((size_t)L);
        for (auto &c : base) c = (uint8_t)(rnd() & 3);
        for (int s = 0; s < ns; ++s) {
            std::vector<uint8_t> q;
            for (int i = 0; i < L; ++i) {
                const uint32_t r = rnd() % 1000;
                if (r < 12) q.push_back((uint8_t)((base[(size_t)i] + 1 + rnd() % 3) & 3));
                else if (r < 16) { q.push_back(base[(size_t)i]); q.push_back((uint8_t)(rnd() & 3)); }
                else if (r < 20) continue;
                else q.push_back(base[(size_t)i]);
            }
            b.push_back(std::move(q));
        }
    }
    poa_b200_engine_t *eng = nullptr;
    if (poa_b200_engine_create(0, nullptr, &eng) != POA_B200_OK) { fprintf(stderr, "engine: %s\n", poa_b200_last_error()); return 3; }
    poa_b200_params_t p = {1, 4, 6, 2, 26, 1, 0, 311, 0.03f, 1, 0};
    // batched reference
    std::vector<int64_t> bso{0}, so{0}; std::vector<int32_t> lens, wts; std::vector<uint8_t> bases;
    for (auto &b : blocks) { for (auto &q : b) { lens.push_back((int32_t)q.size()); wts.push_back(1); bases.insert(bases.end(), q.begin(), q.end()); so.push_back((int64_t)bases.size()); } bso.push_back((int64_t)lens.size()); }
    std::vector<uint64_t> want((size_t)nb), got((size_t)nb);
    double t_batched = 0;
    for (int rep = 0; rep < 2; ++rep) {  // first pass warms the pools
        auto t0 = std::chrono::steady_clock::now();
        poa_b200_result_t *res = nullptr;
        if (poa_b200_run_batch(eng, &p, nb, bso.data(), lens.data(), so.data(), bases.data(), wts.data(), &res) != POA_B200_OK) { fprintf(stderr, "run_batch: %s\n", poa_b200_last_error()); return 3; }
        t_batched = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        for (int b = 0; b < nb; ++b) poa_b200_result_block_hash(res, b, &want[(size_t)b]);
        poa_b200_result_free(res);
    }
    std::atomic<int> failures{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> ths;
    for (int t = 0; t < nt; ++t)
        ths.emplace_back([&, t] {
            std::vector<std::pair<uint64_t, int>> pending;
            auto collect = [&](std::pair<uint64_t, int> tk) {
                poa_b200_result_t *r = nullptr;
                if (poa_b200_wait_block(eng, tk.first, &r) != POA_B200_OK) { ++failures; return; }
                poa_b200_result_block_hash(r, 0, &got[(size_t)tk.second]);
                poa_b200_result_free(r);
            };
            for (int b = t; b < nb; b += nt) {
                std::vector<const uint8_t *> ptr; std::vector<int32_t> ln, w;
                for (auto &q : blocks[(size_t)b]) { ptr.push_back(q.data()); ln.push_back((int32_t)q.size()); w.push_back(1); }
                if (window <= 1) {
                    poa_b200_result_t *r = nullptr;
                    if (poa_b200_poa_block(eng, &p, ns, ptr.data(), ln.data(), w.data(), &r) != POA_B200_OK) { ++failures; continue; }
                    poa_b200_result_block_hash(r, 0, &got[(size_t)b]);
                    poa_b200_result_free(r);
                    continue;
                }
                uint64_t tk = 0;
                if (poa_b200_submit_block(eng, &p, ns, ptr.data(), ln.data(), w.data(), &tk) != POA_B200_OK) { ++failures; continue; }
                pending.emplace_back(tk, b);
                if ((int)pending.size() >= window) { collect(pending.front()); pending.erase(pending.begin()); }
            }
            for (auto &tk : pending) collect(tk);
        });
    for (auto &th : ths) th.join();
    const double t_co = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    int bad = failures.load();
    for (int b = 0; b < nb; ++b) bad += want[(size_t)b] != got[(size_t)b];
    printf("blocks %d threads %d window %d batched_blocks_per_s %.1f coalesced_blocks_per_s %.1f ratio %.3f mismatches %d\n",
           nb, nt, window, nb / t_batched, nb / t_co, t_batched / t_co, bad);
    poa_b200_engine_destroy(eng);
    return bad ? 1 : 0;
}
