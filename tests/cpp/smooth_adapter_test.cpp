// Exercises include/poa_b200_smooth.hpp (the C++ host-side mirror of smooth_abpoa / build_odgi_abPOA, reference
// src/smooth.cpp:133-627, :2442-2574) end to end on the GPU and prints the block graph in a canonical text form that
// tests/test_cpp_adapter.py compares with the same block run through the ctypes binding.
//   usage: smooth_adapter_test <padding> <local 0|1> <consensus-name or -> < sequences (name TAB +|- TAB seq per line)
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include "poa_b200_smooth.hpp"

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage\n"); return 2; }
    const int padding = atoi(argv[1]);
    const bool local = atoi(argv[2]) != 0;
    const std::string cons = std::string(argv[3]) == "-" ? "" : argv[3];
    std::vector<std::string> seqs, names; std::vector<bool> revs;
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream is(line);
        std::string name, strand, seq;
        if (!(is >> name >> strand)) continue;
        is >> seq;  // may be empty
        names.push_back(name); revs.push_back(strand == "-"); seqs.push_back(seq);
    }
    try {
        poa_b200_engine_t *eng = nullptr;
        if (poa_b200_engine_create(0, nullptr, &eng) != POA_B200_OK) { fprintf(stderr, "engine: %s\n", poa_b200_last_error()); return 3; }
        const poa_b200::block_sequences blk = poa_b200::dedup_sequences(seqs, names, revs);  // src/smooth.cpp:217-241
        const poa_b200::block_graph g = poa_b200::smooth_abpoa(eng, blk, 1, 4, 6, 2, 26, 1, padding, local, /*want_msa=*/true, /*banded=*/true, cons);
        printf("dedup %zu", blk.seqs.size());
        for (int w : blk.weights) printf(" %d", w);
        printf("\nnodes %zu\n", g.node_id.size());
        for (size_t i = 0; i < g.node_id.size(); ++i) printf("S %d %c\n", g.node_id[i], g.node_base[i]);
        for (auto &e : g.edges) printf("L %d %d\n", e.first, e.second);
        for (auto &p : g.paths) {
            printf("P %s", p.name.c_str());
            for (auto &s : p.steps) printf(" %d%c", s.node_id, s.is_rev ? '-' : '+');
            printf("\n");
        }
        printf("msa %d %d\n", g.msa_rows, g.msa_len);
        // the graph smooth_abpoa finally returns (unchop + topological order): final_graph_of on the same block
        {
            std::vector<poa_b200::block_sequences> one{blk};
            poa_b200_params_t p = poa_b200::make_params(1, 4, 6, 2, 26, 1, local, true, false, !cons.empty());
            std::vector<int64_t> bso{0}, so{0}; std::vector<int32_t> lens, wts; std::vector<uint8_t> codes;
            for (size_t i = 0; i < blk.seqs.size(); ++i) {
                const size_t at = codes.size(); codes.resize(at + blk.seqs[i].size());
                poa_b200_encode_bases(blk.seqs[i].data(), (int64_t)blk.seqs[i].size(), codes.data() + at);
                lens.push_back((int32_t)blk.seqs[i].size()); wts.push_back(blk.weights[i]); so.push_back((int64_t)codes.size());
            }
            bso.push_back((int64_t)lens.size());
            poa_b200_result_t *res = nullptr;
            if (poa_b200_run_batch(eng, &p, 1, bso.data(), lens.data(), so.data(), codes.data(), wts.data(), &res) != POA_B200_OK) { fprintf(stderr, "run_batch: %s\n", poa_b200_last_error()); return 5; }
            const poa_b200::final_block_graph f = poa_b200::final_graph_of(res, 0, blk, padding, cons, names);
            printf("final %zu\n", f.node_seq.size());
            for (size_t k = 0; k < f.node_seq.size(); ++k) printf("FS %zu %s\n", k + 1, f.node_seq[k].c_str());
            for (auto &e : f.edges) printf("FL %d %d\n", e.first, e.second);
            for (auto &pp : f.paths) {
                printf("FP %s", pp.name.c_str());
                for (auto &st : pp.steps) printf(" %d%c", st.node_id, st.is_rev ? '-' : '+');
                printf("\n");
            }
            poa_b200_result_free(res);
        }
        poa_b200_engine_destroy(eng);
    } catch (const std::exception &e) { fprintf(stderr, "error: %s\n", e.what()); return 4; }
    return 0;
}
