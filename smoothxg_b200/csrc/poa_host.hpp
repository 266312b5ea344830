// poa_host.hpp -- host-side helpers shared by the C ABI (poa_b200.cu) and the debug-only host
// emulation harness (tests/emu): parameter validation, the score matrix, workspace layout.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <string>
#include "../../include/poa_b200.h"
#include "poa_core.cuh"

namespace poa {

inline long long align_up(long long x, long long a) { return (x + a - 1) / a * a; }

int set_err(int code, const std::string &msg);

// abpoa_align.c:12-25 (gen_simple_mat): match on the diagonal, -mismatch elsewhere, 0 for N
inline void build_params(const poa_b200_params_t &p, const poa_b200_engine_opts_t &o, DevParams &d) {
    int match = p.match < 0 ? -p.match : p.match;
    int mismatch = p.mismatch > 0 ? -p.mismatch : p.mismatch;
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) d.mat[i * 5 + j] = i == j ? match : mismatch;
        d.mat[i * 5 + 4] = 0;
    }
    for (int j = 0; j < 5; ++j) d.mat[20 + j] = 0;
    d.o1 = p.gap_open1; d.e1 = p.gap_ext1; d.o2 = p.gap_open2; d.e2 = p.gap_ext2;
    d.oe1 = d.o1 + d.e1; d.oe2 = d.o2 + d.e2;
    d.match = match; d.min_mis = -mismatch;
    d.local = p.align_mode == 1;
    d.wb = p.wb; d.wf = p.wf;
    d.out_cons = p.out_cons ? 1 : 0; d.out_msa = p.out_msa ? 1 : 0;
    // lane counts of the reference build whose vector-granular band-start rule is reproduced (abpoa_align_simd.c:949-960):
    // flags bits 4-5 = 0 AVX-512BW (32 / 16 lanes, what -march=native gives on the hosts this was pinned on), 1 AVX2 (16 / 8),
    // 2 SSE4.1 / NEON (8 / 4).  The survey found results identical across the three on every fixture; the option exists so
    // that a maintainer on a narrower host can match their own build bit for bit should a case ever differ.
    const int isa = (o.flags >> 4) & 3;
    d.pn16 = isa == 1 ? 16 : (isa == 2 ? 8 : 32); d.pn32 = d.pn16 / 2;
    d.emit_cigar = o.emit_cigar ? 1 : 0;
    // packed 16-bit fill (poa_fill16.cuh): every intermediate must stay inside int16, i.e. the slack abPOA
    // builds into inf_min (512 * max(e1,e2), abpoa_align_simd.c:1295) must cover one mismatch, one gap open
    // and the scan's position offsets
    const long long emax = std::max(d.e1, d.e2);
    d.gap_mode = p.gap_open1 == 0 ? 2 : (p.gap_open2 == 0 ? 1 : 0);  // abpoa_align.c:87-91
    {   // packed constants of the 16-bit fill (poa_fill16.cuh); same expressions as the kernel used to evaluate per alignment
        auto pk = [](long long v) { const unsigned h = (unsigned)(v & 0xffff); return h | (h << 16); };
        const long long bm = INT16_MIN;
        const long long inf16 = std::max(bm + d.min_mis, std::max(bm + d.oe1, bm + d.oe2)) + 512 * emax;  // inf_min_of<short>()
        d.pk_inf = pk(inf16); d.pk_negl = pk(-32768 + 8 * emax + 8);
        d.pk_noe1 = pk(-d.oe1); d.pk_noe2 = pk(-d.oe2); d.pk_ne1 = pk(-d.e1); d.pk_ne2 = pk(-d.e2);
        d.pk_ne1_2 = pk(-2LL * d.e1); d.pk_ne1_3 = pk(-3LL * d.e1); d.pk_ne2_2 = pk(-2LL * d.e2); d.pk_ne2_3 = pk(-3LL * d.e2);
        d.pk_ncw1 = pk(-256LL * d.e1); d.pk_ncw2 = pk(-256LL * d.e2);
    }
    d.p16_ok = d.gap_mode == 0 && (o.flags & 1) == 0 && emax >= 1 && emax <= 100 && 240 * emax >= (long long)d.min_mis + std::max(d.oe1, d.oe2) + 64
               && d.oe1 > d.e1 && d.oe2 > d.e2;
    // smoothxg's default scoring: fill_p16<.., PRESET = true> carries these constants as immediates; the literals there are checked
    // against the values computed above, so a change to either side falls back to the generic instantiation instead of diverging
    {
        auto pk = [](long long v) { const unsigned h = (unsigned)(v & 0xffff); return h | (h << 16); };
        d.p16_default = d.p16_ok && match == 1 && mismatch == -4 && d.o1 == 6 && d.e1 == 2 && d.o2 == 26 && d.e2 == 1
                        && d.pk_inf == pk(-31717) && d.pk_negl == pk(-32744) && d.pk_noe1 == pk(-8) && d.pk_noe2 == pk(-27) && d.pk_ne1 == pk(-2)
                        && d.pk_ne2 == pk(-1) && d.pk_ne1_2 == pk(-4) && d.pk_ne1_3 == pk(-6) && d.pk_ne2_2 == pk(-2) && d.pk_ne2_3 == pk(-3)
                        && d.pk_ncw1 == pk(-512) && d.pk_ncw2 == pk(-256);
    }
}

inline int check_params(const poa_b200_params_t &p) {
    if (p.gap_open1 < 0 || p.gap_open2 < 0 || p.gap_ext1 < 0 || p.gap_ext2 < 0 || p.align_mode < 0 || p.align_mode > 1)
        return set_err(POA_B200_EARG, "negative gap penalty or bad align_mode");
    if (p.gap_ext1 == 0 && p.gap_ext2 == 0)
        return set_err(POA_B200_EUNSUP, "gap extension 0: the reference's 16-bit scores have no head room (inf_min = INT16_MIN + ...); refused");
    // The kernel evaluates F as a max-plus prefix scan in 32-bit registers; that equals the reference's
    // wrapping 16-bit arithmetic as long as no junk cell can fall below INT16_MIN, which abPOA's own
    // inf_min margin guarantees for any sane scoring (abpoa_align_simd.c:1295).
    long long mm = p.mismatch < 0 ? -(long long)p.mismatch : p.mismatch;
    long long oe1 = (long long)p.gap_open1 + p.gap_ext1, oe2 = (long long)p.gap_open2 + p.gap_ext2;
    long long margin = std::max(mm, std::max(oe1, oe2)) + 512LL * std::max(p.gap_ext1, p.gap_ext2);
    if (margin < mm + std::max(oe1, oe2) + std::max(p.gap_ext1, p.gap_ext2))
        return set_err(POA_B200_EUNSUP, "scoring parameters let 16-bit scores underflow in the reference; refused");
    if (mm > 10000 || oe1 > 10000 || oe2 > 10000 || std::abs((long long)p.match) > 10000)
        return set_err(POA_B200_EUNSUP, "scoring parameters out of the supported range");
    return POA_B200_OK;
}

// Lay the per-CTA workspace out for blocks of at most `nmax` nodes, `max_bases` total bases,
// `max_len` bases per sequence and `max_seq` sequences.
inline void make_layout(WsLayout &L, long long nmax, long long max_bases, long long max_len, long long max_seq,
                 long long pool_growth, long long slab_bytes, int emit_cigar) {
    long long o = 0;
    auto take = [&](long long bytes) { long long at = o; o = align_up(o + bytes, 256); return at; };
    L.nmax = (int)nmax;
    L.pool_cap = (int)(4 * nmax + pool_growth);
    const long long pc = L.pool_cap;
    L.o_base = take(nmax); L.o_aln_n = take(nmax); L.o_aln = take(16 * nmax);
    L.o_in_off = take(4 * nmax); L.o_in_n = take(4 * nmax); L.o_out_off = take(4 * nmax); L.o_out_n = take(4 * nmax);
    L.o_pool_id = take(4 * pc); L.o_pool_w = take(4 * pc); L.o_pool_row = take(4 * pc);
    L.o_idx2id = take(4 * nmax); L.o_id2idx = take(4 * nmax); L.o_remain = take(4 * nmax);
    L.o_tmp0 = take(4 * nmax); L.o_tmp1 = take(4 * nmax + 64); L.o_tmp2 = take(4 * nmax); L.o_tmp3 = take(4 * nmax + 64);
    L.o_rowinfo = take(16 * nmax); L.o_rowmeta = take(16 * nmax); L.o_rbase = take(nmax);
    L.o_rr = take(4 * nmax); L.o_mplr = take(4 * nmax); L.o_mprr = take(4 * nmax); L.o_pred4 = take(16 * nmax);
    long long cig_one = max_len + nmax + 8;
    L.cig_cap = (int)std::min<long long>(emit_cigar ? cig_one * std::max<long long>(max_seq, 1) : cig_one, INT32_MAX);
    L.o_cig = take(8LL * L.cig_cap);
    L.o_path = take(4 * std::max<long long>(max_bases, 1));
    L.o_best = take(4 * std::max<long long>(max_seq, 1)); L.o_ncig = take(4 * std::max<long long>(max_seq, 1)); L.o_plen = take(4 * std::max<long long>(max_seq, 1));
    L.o_nrun = take(8 * (std::max<long long>(max_seq, 1) + 2));
    L.o_qp = take(5 * ((max_len >> 8) + 2) * 512);
    L.slab_bytes = align_up(slab_bytes, 512);
    L.slab_planes = (unsigned)std::min<long long>(L.slab_bytes / 512, 0xfff00000ll);
    L.o_slab = take(L.slab_bytes);
    L.stride = o;
}


}  // namespace poa
