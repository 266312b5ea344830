// poa_wire.hpp -- host-side decoder of a block's result body (the wire format poa_block() writes, see WireLayout in
// poa_core.cuh) into the flat int32 arrays poa_b200_block_view_t exposes.  Plain C++, no CUDA: used by the C ABI
// (poa_b200.cu) and by the CPU emulation harness of the device code (tests/emu).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include "poa_core.cuh"

namespace poa {

// One decoded block: every int32 section lives in `buf`; cigar words and MSA bytes are read in place from the arena.
struct DecodedBlock {
    std::vector<int32_t> buf;
    long long base = 0, in_n = 0, in_id = 0, in_w = 0, out_n = 0, out_id = 0, out_w = 0, aln_n = 0, aln_id = 0;  // offsets into buf
    long long path_len = 0, path_node = 0, cons_node = 0, best = 0, ncig = 0;
    const int32_t *cig = nullptr;
    const uint8_t *msa = nullptr;
};

inline bool wire_header_ok(const int *h) {
    if (h[H_FORMAT] != WIRE_NARROW && h[H_FORMAT] != WIRE_WIDE) return false;
    if (h[H_N_NODE] < 0 || h[H_N_SEQ] < 0 || h[H_IN_TOT] < 0 || h[H_OUT_TOT] < 0 || h[H_ALN_TOT] < 0 || h[H_PATH_TOT] < 0 || h[H_CIG_TOT] < 0 || h[H_RUN_TOT] < 0) return false;
    WireLayout W;
    const long long msa_bytes = (long long)h[H_MSA_ROWS] * (h[H_MSA_LEN] > 0 ? h[H_MSA_LEN] : 0);
    wire_layout(W, h[H_N_NODE], h[H_N_SEQ], h[H_IN_TOT], h[H_OUT_TOT], h[H_ALN_TOT], h[H_RUN_TOT], h[H_CIG_TOT], msa_bytes, h[H_FORMAT]);
    return W.words == (long long)h[H_BODY_WORDS];
}

// body = first arena word of the block.  Returns false if the run lists are inconsistent with the header's totals.
inline bool wire_decode(const int *h, const int *body, DecodedBlock &d) {
    const int n = h[H_N_NODE], ns = h[H_N_SEQ], format = h[H_FORMAT];
    const long long in_tot = h[H_IN_TOT], out_tot = h[H_OUT_TOT], aln_tot = h[H_ALN_TOT], path_tot = h[H_PATH_TOT], run_tot = h[H_RUN_TOT];
    const int cons_len = h[H_CONS_LEN] > 0 ? h[H_CONS_LEN] : 0;
    const long long msa_bytes = (long long)h[H_MSA_ROWS] * (h[H_MSA_LEN] > 0 ? h[H_MSA_LEN] : 0);
    WireLayout W;
    wire_layout(W, n, ns, in_tot, out_tot, aln_tot, run_tot, h[H_CIG_TOT], msa_bytes, format);
    const bool wide = format == WIRE_WIDE;
    long long o = 0;
    auto take = [&o](long long k) { const long long at = o; o += k; return at; };
    d.base = take(n); d.in_n = take(n); d.in_id = take(in_tot); d.in_w = take(in_tot);
    d.out_n = take(n); d.out_id = take(out_tot); d.out_w = take(out_tot); d.aln_n = take(n); d.aln_id = take(aln_tot);
    d.path_len = take(ns); d.path_node = take(path_tot); d.cons_node = take(cons_len); d.best = take(ns); d.ncig = take(ns);
    d.buf.resize((size_t)o);
    int32_t *b = d.buf.data();
    auto bytes = [&](long long sec, long long dst, long long k) { const uint8_t *p = (const uint8_t *)(body + sec); for (long long i = 0; i < k; ++i) b[dst + i] = p[i]; };
    auto elems = [&](long long sec, long long dst, long long k) {
        if (wide) memcpy(b + dst, body + sec, (size_t)k * 4);
        else { const uint16_t *p = (const uint16_t *)(body + sec); for (long long i = 0; i < k; ++i) b[dst + i] = p[i]; }
    };
    bytes(W.o_base, d.base, n); bytes(W.o_aln_n, d.aln_n, n);
    elems(W.o_in_n, d.in_n, n); elems(W.o_out_n, d.out_n, n);
    elems(W.o_in_id, d.in_id, in_tot); elems(W.o_in_w, d.in_w, in_tot);
    elems(W.o_out_id, d.out_id, out_tot); elems(W.o_out_w, d.out_w, out_tot); elems(W.o_aln_id, d.aln_id, aln_tot);
    memcpy(b + d.path_len, body + W.o_plen, (size_t)ns * 4);
    memcpy(b + d.best, body + W.o_best, (size_t)ns * 4);
    memcpy(b + d.ncig, body + W.o_ncig, (size_t)ns * 4);
    // runs -> one node id per path step
    const int32_t *nrun = body + W.o_nrun;
    auto run_at = [&](long long r, int which) -> long long {
        return wide ? (long long)(uint32_t)body[W.o_runs + 2 * r + which] : (long long)((const uint16_t *)(body + W.o_runs))[2 * r + which];
    };
    long long r = 0, at = d.path_node;
    for (int k = 0; k <= ns; ++k) {
        const long long len = k < ns ? b[d.path_len + k] : cons_len;
        if (len < 0) return false;
        if (k == ns) { if (at != d.path_node + path_tot) return false; at = d.cons_node; }
        const long long nr = nrun[k];
        if (nr < 0 || r + nr > run_tot || (len > 0) != (nr > 0)) return false;
        for (long long x = 0; x < nr; ++x) {
            const long long start = run_at(r + x, 0), pos = run_at(r + x, 1);
            const long long end = x + 1 < nr ? run_at(r + x + 1, 1) : len;
            if (pos < 0 || end <= pos || end > len || (x == 0 && pos != 0)) return false;
            for (long long t = pos; t < end; ++t) b[at + t] = (int32_t)(start + (t - pos));
        }
        r += nr;
        if (k < ns) { at += len; if (at > d.path_node + path_tot) return false; }
    }
    if (r != run_tot) return false;
    d.cig = body + W.o_cig;
    d.msa = (const uint8_t *)(body + W.o_msa);
    return true;
}

}  // namespace poa
