// poa_b200_smooth.hpp -- host-side mirror of smoothxg's per-block abPOA adapter above the C ABI (header only).
//
// Reference: smooth_abpoa (src/smooth.cpp:133-627) and build_odgi_abPOA (:2442-2574).  The reference function
// (a) walks the block's path ranges in XG and produces one oriented, padded string per range (:177-214),
// (b) de-duplicates identical strings, counting multiplicities (:217-241),
// (c) runs abPOA (:256-351), (d) builds the odgi graph with one path per *name* (:2513-2532) and the consensus path.
// Step (a) needs smoothxg's XG index and stays in smoothxg; this header mirrors (b), (c) and (d) with the same
// names and argument meaning, on plain C++ containers, so that the patched smooth_abpoa is a handful of lines
//  and the batched variant can be called once per chunk of blocks.  (c) is the GPU engine behind
// include/poa_b200.h; there is no CPU path.  Errors: the reference exits the process (err_fatal, exit(1) :350,
// :553); here they are thrown as std::runtime_error.
#ifndef POA_B200_SMOOTH_HPP
#define POA_B200_SMOOTH_HPP

#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "mash_b200.h"
#include "poa_b200.h"

namespace poa_b200 {

// What smooth_abpoa holds after the de-duplication loop (src/smooth.cpp:217-241).
struct block_sequences {
    std::vector<std::string> seqs;                          // unique sequences, first-occurrence order
    std::vector<int32_t> weights;                           // multiplicity of each (weights[rank]++, :236)
    std::vector<std::vector<std::string>> dup_seq_names;    // names sharing each sequence
    std::vector<std::vector<bool>> dup_is_revs;             // their orientation flags
};

// ---- a2: one path range -> the string handed to the POA (src/smooth.cpp:75-126 append_to_sequence, :177-214).
// Graph is any type with the handle-graph calls the reference makes on xg::XG: get_path_handle_of_step, path_begin, path_end,
// get_handle_of_step, get_length, get_sequence (oriented), get_is_reverse, get_previous_step, get_next_step; Step is its step handle.
// The reference's behaviour is the specification here, including what looks accidental: the left flank starts AT the range's first
// step (so it repeats bases of that node), collects node sequences in walk order (backwards along the path, each node spelled
// forwards) and never visits the path's first step; both flanks take the LAST `need` bases of a node that is longer than what is
// still needed; flanks that run off the path are filled with 'N' (left: in front, right: behind).  `rev_comp` is the caller's
// reverse-complement (the reference uses odgi::reverse_complement_in_place); it is applied when more bases were walked on the
// reverse strand than on the forward strand (:208-210).  Returns the oriented, padded string; *is_rev tells the orientation.
template <class Graph, class Step, class RevComp>
inline std::string extract_range_sequence(const Graph &graph, const Step &range_begin, const Step &range_end, int poa_padding,
                                          RevComp rev_comp, bool *is_rev = nullptr) {
    std::string seq;
    uint64_t fwd_bp = 0, rev_bp = 0;
    const auto path = graph.get_path_handle_of_step(range_begin);
    auto flank = [&](const Step &from, bool left) {           // :75-126
        Step step = from;
        const Step stop = left ? graph.path_begin(path) : graph.path_end(path);
        uint64_t need = (uint64_t)poa_padding;
        std::string got;
        while (step != stop && need > 0) {
            const auto h = graph.get_handle_of_step(step);
            const uint64_t l = graph.get_length(h);
            const std::string node_seq = graph.get_sequence(h);
            const uint64_t take = l <= need ? l : need;
            got.append(l <= need ? node_seq : node_seq.substr(node_seq.size() - need));
            (graph.get_is_reverse(h) ? rev_bp : fwd_bp) += take;
            need -= take;
            step = left ? graph.get_previous_step(step) : graph.get_next_step(step);
        }
        if (left) { seq.append(need, 'N'); seq.append(got); } else { seq.append(got); seq.append(need, 'N'); }
    };
    flank(range_begin, true);
    for (Step step = range_begin; step != range_end; step = graph.get_next_step(step)) {   // :191-201
        const auto h = graph.get_handle_of_step(step);
        seq.append(graph.get_sequence(h));
        (graph.get_is_reverse(h) ? rev_bp : fwd_bp) += graph.get_length(h);
    }
    flank(range_end, false);
    const bool rev = rev_bp > fwd_bp;                         // :208-210
    if (rev) rev_comp(seq);
    if (is_rev) *is_rev = rev;
    return seq;
}

// src/smooth.cpp:217-241.  The reference keys on XXH64(seq) only; identical strings always share a key, so keying
// on the string itself gives the same grouping except on a 64-bit hash collision between different strings.
inline block_sequences dedup_sequences(const std::vector<std::string> &seqs, const std::vector<std::string> &names,
                                       const std::vector<bool> &is_revs) {
    if (seqs.size() != names.size() || seqs.size() != is_revs.size()) throw std::runtime_error("poa_b200: ragged block input");
    block_sequences b;
    std::unordered_map<std::string, size_t> rank_of;
    for (size_t i = 0; i < seqs.size(); ++i) {
        auto it = rank_of.find(seqs[i]);
        if (it == rank_of.end()) {
            rank_of.emplace(seqs[i], b.seqs.size());
            b.seqs.push_back(seqs[i]); b.weights.push_back(1);
            b.dup_seq_names.push_back({names[i]}); b.dup_is_revs.push_back({is_revs[i]});
        } else {
            b.weights[it->second]++;
            b.dup_seq_names[it->second].push_back(names[i]); b.dup_is_revs[it->second].push_back(is_revs[i]);
        }
    }
    return b;
}

// Score preset smooth_and_lace picks from the block's estimated identity when -a is given (src/smooth.cpp:2028-2062;
// the MinHash estimate itself, rkmh::hash_sequences / rkmh::compare :2005-2023, is mash_b200_block_identity(), include/mash_b200.h).  Returns false and
// leaves the scores alone below 0.90 ("use the set/default penalties").  Blocks with different presets go into one
// batch per preset (the engine's parameters are per batch).
inline bool adaptive_poa_preset(float est_identity_threshold, int &poa_m, int &poa_n, int &poa_g, int &poa_e, int &poa_q, int &poa_c) {
    // the reference compares the float with double literals (0.95f and 0.9f are below 0.95 and 0.9); same table as
    // mash_b200_preset() (include/mash_b200.h), which also computes the estimate itself on the GPU
    static const double thr[5] = {0.99, 0.98, 0.97, 0.95, 0.90};
    static const int presets[5][6] = {{1, 19, 39, 3, 81, 1}, {1, 13, 31, 3, 51, 1}, {1, 9, 16, 2, 41, 1}, {1, 7, 11, 2, 33, 1}, {1, 4, 6, 2, 26, 1}};
    for (int r = 0; r < 5; ++r)
        if ((double)est_identity_threshold >= thr[r]) {
            const int *p = presets[r];
            poa_m = p[0]; poa_n = p[1]; poa_g = p[2]; poa_e = p[3]; poa_q = p[4]; poa_c = p[5];
            return true;
        }
    return false;
}

// Batched form of the estimate itself (src/smooth.cpp:1982-2023): `ranges[b]` = the raw strings of block b's path ranges
// (the XG walk :1987-1992, before padding / orientation / de-duplication), upper case.  Returns est_identity_threshold per
// block, -1 where fewer than two strings have 8*kmer_size bases (the reference then keeps the user's scores).  Runs on the
// GPU (include/mash_b200.h); blocks deeper than max_block_depth_for_padding_more (:1984) are the caller's to leave out.
inline std::vector<float> estimate_block_identity(int device, const std::vector<std::vector<std::string>> &ranges, int kmer_size = 17) {
    std::vector<int64_t> block_seq_off(1, 0), seq_off(1, 0);
    std::vector<int32_t> seq_len;
    std::string bases;
    for (auto &blk : ranges) {
        for (auto &s : blk) { bases += s; seq_len.push_back((int32_t)s.size()); seq_off.push_back((int64_t)bases.size()); }
        block_seq_off.push_back((int64_t)seq_len.size());
    }
    std::vector<float> thr(ranges.size(), -1.0f);
    if (ranges.empty()) return thr;
    if (bases.empty()) bases.push_back('N');
    const int rc = mash_b200_block_identity(device, kmer_size, (int64_t)ranges.size(), block_seq_off.data(), seq_len.data(), seq_off.data(),
                                            bases.data(), thr.data(), nullptr, nullptr, nullptr, nullptr);
    if (rc != MASH_B200_OK) throw std::runtime_error(std::string("poa_b200: identity estimate failed: ") + mash_b200_last_error());
    return thr;
}

struct step_t { int32_t node_id; bool is_rev; };            // odgi handle: id + orientation
struct path_t { std::string name; std::vector<step_t> steps; };

// What build_odgi_abPOA leaves in `output` (src/smooth.cpp:2442-2574), before unchop/sort (:545-620).
struct block_graph {
    std::vector<int32_t> node_id;                           // creation order
    std::string node_base;                                  // one base per node
    std::vector<std::pair<int32_t, int32_t>> edges;         // forward-forward
    std::vector<path_t> paths;                              // one per name, consensus last when requested
    int32_t msa_len = -1, msa_rows = 0;
    std::vector<uint8_t> msa;                               // abc->msa_base rows (gap = 5), when asked for
};

inline poa_b200_params_t make_params(int poa_m, int poa_n, int poa_g, int poa_e, int poa_q, int poa_c,
                                     bool local_alignment, bool banded_alignment, bool want_msa, bool add_consensus) {
    poa_b200_params_t p;                                    // src/smooth.cpp:256-297
    p.match = poa_m; p.mismatch = poa_n; p.gap_open1 = poa_g; p.gap_ext1 = poa_e; p.gap_open2 = poa_q; p.gap_ext2 = poa_c;
    p.align_mode = local_alignment ? 1 : 0;
    p.wb = banded_alignment ? 311 : -1; p.wf = 0.03f;
    p.out_cons = add_consensus ? 1 : 0; p.out_msa = want_msa ? 1 : 0;
    return p;
}

// What smooth_abpoa returns (src/smooth.cpp:545-620): the block graph after odgi unchop, renumbered 1..n in a topological order,
// one edge per consecutive pair of path steps, paths in the block's original order with the consensus last.
struct final_block_graph {
    std::vector<std::string> node_seq;                      // node k (0-based) has id k + 1
    std::vector<std::pair<int32_t, int32_t>> edges;         // forward-forward, ids
    std::vector<path_t> paths;
};

namespace detail {
inline void check(int rc, const char *what) {
    if (rc != POA_B200_OK) throw std::runtime_error(std::string("poa_b200: ") + what + ": " + poa_b200_strerror(rc) + ": " + poa_b200_last_error());
}

inline block_graph graph_of(const poa_b200_result_t *res, int64_t blk, const block_sequences &b, int padding_len,
                            const std::string &consensus_name) {
    poa_b200_block_view_t v;
    check(poa_b200_result_block(res, blk, &v), "result_block");
    if (v.status != POA_B200_OK) throw std::runtime_error(std::string("poa_b200: block failed: ") + poa_b200_strerror(v.status));
    const bool add_consensus = !consensus_name.empty();
    if (add_consensus && v.cons_len < 0) throw std::runtime_error("poa_b200: no consensus sequence generated");  // :348-351
    poa_b200_graph_t *g = nullptr;
    check(poa_b200_block_graph(&v, padding_len, add_consensus ? 1 : 0, &g), "block_graph");
    poa_b200_graph_view_t gv;
    check(poa_b200_graph_view(g, &gv), "graph_view");
    block_graph out;
    out.node_id.assign(gv.node_id, gv.node_id + gv.n_node);
    out.node_base.assign(gv.node_base, gv.node_base + gv.n_node);
    for (int32_t i = 0; i < gv.n_edge; ++i) out.edges.emplace_back(gv.edge_from[i], gv.edge_to[i]);
    for (size_t i = 0; i < b.seqs.size() && (int32_t)i < gv.n_path; ++i) {  // :2513-2532: one path per name
        const int32_t *s0 = gv.path_node + gv.path_off[i], *s1 = gv.path_node + gv.path_off[i + 1];
        for (size_t z = 0; z < b.dup_seq_names[i].size(); ++z) {
            path_t p; p.name = b.dup_seq_names[i][z];
            if (b.dup_is_revs[i][z]) for (const int32_t *s = s1; s != s0;) p.steps.push_back({*--s, true});
            else for (const int32_t *s = s0; s != s1; ++s) p.steps.push_back({*s, false});
            out.paths.push_back(std::move(p));
        }
    }
    if (add_consensus) {                                     // :2534-2549
        path_t p; p.name = consensus_name;
        for (int64_t k = gv.path_off[gv.n_path - 1]; k < gv.path_off[gv.n_path]; ++k) p.steps.push_back({gv.path_node[k], false});
        out.paths.push_back(std::move(p));
    }
    if (v.msa_len >= 0) { out.msa_len = v.msa_len; out.msa_rows = v.msa_rows; out.msa.assign(v.msa, v.msa + (size_t)v.msa_rows * v.msa_len); }
    poa_b200_graph_free(g);
    return out;
}
}  // namespace detail

// The same block as smooth_abpoa finally returns it (poa_b200_block_final_graph): `names_in_original_order` is the block's
// all_names_in_original_order (:242, one name per path range, duplicates of a sequence included); paths come out in that order.
inline final_block_graph final_graph_of(const poa_b200_result_t *res, int64_t blk, const block_sequences &b, int padding_len,
                                        const std::string &consensus_name, const std::vector<std::string> &names_in_original_order) {
    poa_b200_block_view_t v;
    detail::check(poa_b200_result_block(res, blk, &v), "result_block");
    if (v.status != POA_B200_OK) throw std::runtime_error(std::string("poa_b200: block failed: ") + poa_b200_strerror(v.status));
    const bool add_consensus = !consensus_name.empty();
    poa_b200_graph_t *g = nullptr;
    detail::check(poa_b200_block_final_graph(&v, padding_len, add_consensus ? 1 : 0, &g), "block_final_graph");
    poa_b200_final_graph_view_t gv;
    detail::check(poa_b200_final_graph_view(g, &gv), "final_graph_view");
    final_block_graph out;
    for (int32_t k = 0; k < gv.n_node; ++k) out.node_seq.emplace_back(gv.seq + gv.seq_off[k], gv.seq + gv.seq_off[k + 1]);
    for (int32_t i = 0; i < gv.n_edge; ++i) out.edges.emplace_back(gv.edge_from[i], gv.edge_to[i]);
    std::unordered_map<std::string, std::pair<size_t, bool>> owner;  // name -> (deduplicated sequence, reverse strand)
    for (size_t i = 0; i < b.seqs.size(); ++i)
        for (size_t z = 0; z < b.dup_seq_names[i].size(); ++z) owner[b.dup_seq_names[i][z]] = {i, b.dup_is_revs[i][z]};
    for (auto &name : names_in_original_order) {             // :608-620
        auto it = owner.find(name);
        if (it == owner.end()) throw std::runtime_error("poa_b200: unknown path name " + name);
        const size_t i = it->second.first;
        path_t p; p.name = name;
        const int32_t *s0 = gv.path_node + gv.path_off[i], *s1 = gv.path_node + gv.path_off[i + 1];
        if (it->second.second) for (const int32_t *s = s1; s != s0;) p.steps.push_back({*--s, true});
        else for (const int32_t *s = s0; s != s1; ++s) p.steps.push_back({*s, false});
        out.paths.push_back(std::move(p));
    }
    if (add_consensus) {
        path_t p; p.name = consensus_name;
        for (int64_t k = gv.path_off[gv.n_path - 1]; k < gv.path_off[gv.n_path]; ++k) p.steps.push_back({gv.path_node[k], false});
        out.paths.push_back(std::move(p));
    }
    poa_b200_graph_free(g);
    poa_b200_result_release_block(res, blk);
    return out;
}

// Batched form of smooth_abpoa (src/smooth.cpp:133-627) from the de-duplicated sequences on: one GPU launch for all
// blocks.  `padding_len[b]` is the block's poa_padding (:1946-1970), `consensus_name[b]` empty = no consensus path.
inline std::vector<block_graph> smooth_abpoa_batch(poa_b200_engine_t *engine, const std::vector<block_sequences> &blocks,
                                                   int poa_m, int poa_n, int poa_g, int poa_e, int poa_q, int poa_c,
                                                   const std::vector<int> &padding_len, bool local_alignment, bool want_msa,
                                                   const std::vector<std::string> &consensus_name, bool banded_alignment = true) {
    const size_t nb = blocks.size();
    if (padding_len.size() != nb || consensus_name.size() != nb) throw std::runtime_error("poa_b200: ragged batch input");
    bool any_cons = false;
    for (auto &c : consensus_name) any_cons |= !c.empty();
    const poa_b200_params_t params = make_params(poa_m, poa_n, poa_g, poa_e, poa_q, poa_c, local_alignment, banded_alignment, want_msa, any_cons);
    std::vector<int64_t> block_seq_off(1, 0), seq_off(1, 0);
    std::vector<int32_t> seq_len, weight;
    std::vector<uint8_t> bases;
    for (auto &b : blocks) {
        for (size_t i = 0; i < b.seqs.size(); ++i) {
            const size_t at = bases.size();
            bases.resize(at + b.seqs[i].size());
            poa_b200_encode_bases(b.seqs[i].data(), (int64_t)b.seqs[i].size(), bases.data() + at);  // :304-313
            seq_len.push_back((int32_t)b.seqs[i].size()); weight.push_back(b.weights[i]);
            seq_off.push_back((int64_t)bases.size());
        }
        block_seq_off.push_back((int64_t)seq_len.size());
    }
    poa_b200_result_t *res = nullptr;
    const int rc = poa_b200_run_batch(engine, &params, (int64_t)nb, block_seq_off.data(), seq_len.data(), seq_off.data(), bases.data(), weight.data(), &res);
    if (rc != POA_B200_OK && rc != POA_B200_EBLOCK) detail::check(rc, "run_batch");
    std::vector<block_graph> out;
    try {
        for (size_t b = 0; b < nb; ++b) out.push_back(detail::graph_of(res, (int64_t)b, blocks[b], padding_len[b], consensus_name[b]));
    } catch (...) { poa_b200_result_free(res); throw; }
    poa_b200_result_free(res);
    return out;
}

// Per-block form with smooth_abpoa's parameter list (src/smooth.cpp:133-150) minus the XG arguments.
inline block_graph smooth_abpoa(poa_b200_engine_t *engine, const block_sequences &block,
                                int poa_m, int poa_n, int poa_g, int poa_e, int poa_q, int poa_c,
                                int poa_padding, bool local_alignment, bool want_msa, bool banded_alignment,
                                const std::string &consensus_name) {
    return smooth_abpoa_batch(engine, {block}, poa_m, poa_n, poa_g, poa_e, poa_q, poa_c, {poa_padding}, local_alignment, want_msa,
                              {consensus_name}, banded_alignment)[0];
}

}  // namespace poa_b200
#endif  // POA_B200_SMOOTH_HPP
