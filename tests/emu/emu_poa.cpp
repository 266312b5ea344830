// TEST INFRASTRUCTURE ONLY -- debug harness, never built by __graft_entry__.build(), never shipped,
// never loaded by the smoothxg_b200 package.
//
// Compiles smoothxg_b200/csrc/poa_core.cuh as plain C++ with -DPOA_HOST_EMU (one emulated thread,
// warp size 1) and runs poa_block<1>() for one block on the host, so that the serial device logic
// (vector indexing, band bookkeeping, traceback, fusion, topological sort, output packing) can be
// checked against the oracle on a machine without a GPU.  The warp/block collectives degenerate to
// identities here; they are only exercised by the -m gpu tests.
//
// Output: the canonical dump of oracle/poa_dump.h, so tests compare with np.array_equal.
#define POA_HOST_EMU
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../smoothxg_b200/csrc/poa_host.hpp"
#include "../../smoothxg_b200/csrc/poa_wire.hpp"
#include "../../oracle/poa_dump.h"

namespace poa {
int set_err(int code, const std::string &msg) { fprintf(stderr, "[emu] %s\n", msg.c_str()); return code; }
}
using namespace poa;

#if POA_EMU_LANES > 1
// ---- POA_EMU_NW warps of 32 lock-step lanes as fibers: a warp collective publishes the lane's value and yields
// until the other 31 lanes of its warp arrived; a block barrier waits for every fiber.
#ifndef POA_EMU_NW
#define POA_EMU_NW 1
#endif
#include <ucontext.h>
#include <functional>
namespace poa_emu {
static const int NL = 32 * POA_EMU_NW;
static ucontext_t g_main, g_ctx[NL];
static int g_cur = 0, g_done[NL];
static long long g_gen[NL], g_arrivals[POA_EMU_NW], g_bgen[NL], g_barrivals = 0, g_progress = 0;
static int g_buf[POA_EMU_NW][2][32];
static std::function<void()> g_body;
int lane() { return g_cur; }
static void yield_to_main() { swapcontext(&g_ctx[g_cur], &g_main); }
void xchg(int v, int *all) {
    const int me = g_cur, w = me >> 5, l = me & 31;
    const long long gen = g_gen[me]++;
    g_buf[w][gen & 1][l] = v;
    ++g_arrivals[w]; ++g_progress;
    while (g_arrivals[w] < 32LL * (gen + 1)) yield_to_main();
    for (int i = 0; i < 32; ++i) all[i] = g_buf[w][gen & 1][i];
}
void sync_all() {
    const int me = g_cur;
    const long long gen = g_bgen[me]++;
    ++g_barrivals; ++g_progress;
    while (g_barrivals < (long long)NL * (gen + 1)) yield_to_main();
}
static void trampoline() { g_body(); g_done[g_cur] = 1; ++g_progress; yield_to_main(); }
static void run_warp(std::function<void()> body) {
    static std::vector<char> stacks;
    const size_t SS = 1 << 20;
    stacks.assign(SS * NL, 0);
    g_body = body; g_barrivals = 0; g_progress = 0;
    for (int w = 0; w < POA_EMU_NW; ++w) g_arrivals[w] = 0;
    for (int i = 0; i < NL; ++i) {
        g_done[i] = 0; g_gen[i] = 0; g_bgen[i] = 0;
        getcontext(&g_ctx[i]);
        g_ctx[i].uc_stack.ss_sp = stacks.data() + SS * i; g_ctx[i].uc_stack.ss_size = SS; g_ctx[i].uc_link = &g_main;
        makecontext(&g_ctx[i], trampoline, 0);
    }
    for (;;) {
        int alive = 0;
        const long long before = g_progress;
        for (int i = 0; i < NL; ++i) if (!g_done[i]) { ++alive; g_cur = i; swapcontext(&g_main, &g_ctx[i]); }
        if (!alive) break;
        if (g_progress == before) { fprintf(stderr, "[emu] divergent collective: lanes deadlocked\n"); abort(); }
    }
}
}  // namespace poa_emu
#endif

extern "C" void emu_free(void *p) { free(p); }

// Same run, but returns the product's wire format: HDR_WORDS header words followed by the arena words
// (so the C ABI's result accessors and poa_b200_result_from_parts can be exercised without a GPU).
static int g_want_wire = 0;
extern "C" int32_t *emu_poa_block(const pd_params_t *pp, int n_seq, const int32_t *seq_len, const uint8_t *bases,
                                  const int32_t *weight, int instrument, int64_t *n_out);
extern "C" int32_t *emu_poa_block_wire(const pd_params_t *pp, int n_seq, const int32_t *seq_len, const uint8_t *bases,
                                       const int32_t *weight, int instrument, int64_t *n_out) {
    g_want_wire = 1;
    int32_t *r = emu_poa_block(pp, n_seq, seq_len, bases, weight, instrument, n_out);
    g_want_wire = 0;
    return r;
}

extern "C" int32_t *emu_poa_block(const pd_params_t *pp, int n_seq, const int32_t *seq_len, const uint8_t *bases,
                                  const int32_t *weight, int instrument, int64_t *n_out) {
    *n_out = 0;
    poa_b200_params_t p;
    static_assert(sizeof(pd_params_t) == sizeof(poa_b200_params_t), "param structs must match");
    memcpy(&p, pp, sizeof(p));
    if (check_params(p)) return nullptr;
    poa_b200_engine_opts_t opts; memset(&opts, 0, sizeof(opts));
    opts.emit_cigar = instrument ? 1 : 0;
    if (getenv("POA_EMU_NO_P16")) opts.flags = 1;
    DevParams dp; build_params(p, opts, dp);
    long long tot = 0, max_len = 1;
    for (int i = 0; i < n_seq; ++i) { tot += seq_len[i]; if (seq_len[i] > max_len) max_len = seq_len[i]; }
    const char *lvl = getenv("POA_EMU_TIGHT");
    long long nmax = std::max<long long>(std::max<long long>(tot + 2, n_seq + 2), 1024);
    long long slab = nmax * std::max<long long>(((max_len + 1) / 8 + 2) * 5 * 32, ((max_len + 1) / 256 + 2) * 2560);
    long long growth = 8 * (tot + n_seq) + 64;
    if (lvl) { nmax = std::max<long long>(atoll(lvl), max_len + 2); slab = slab / 64; growth = 64; }
    WsLayout L;
    make_layout(L, nmax, std::max<long long>(tot, 1), max_len, std::max(n_seq, 1), growth, slab, dp.emit_cigar);
    std::vector<char> ws((size_t)L.stride + 256);
    char *wsp = (char *)(((uintptr_t)ws.data() + 255) & ~(uintptr_t)255);
    std::vector<long long> bso{0, n_seq}, so((size_t)n_seq + 1, 0);
    for (int i = 0; i < n_seq; ++i) so[(size_t)i + 1] = so[(size_t)i] + seq_len[i];
    int order0 = 0;
    DevBatch B; B.block_seq_off = bso.data(); B.seq_len = seq_len; B.seq_off = so.data(); B.bases = bases; B.weight = weight; B.order = &order0; B.n_order = 1;
    std::vector<int> hdr(HDR_WORDS, -1);
    long long cap = 16 * nmax + tot + 64 + (long long)(n_seq + 1) * nmax / 4 + 2 * (tot + (long long)n_seq * (nmax + max_len + 8));
    std::vector<int> arena((size_t)cap);
    unsigned long long used = 0, phase[PH_N] = {0};
    int counter = 0;
    DevOut O; O.hdr = hdr.data(); O.arena = arena.data(); O.arena_used = &used; O.arena_cap = (unsigned long long)cap; O.phase = phase; O.counter = &counter;
    Shared sh; memset(&sh, 0, sizeof(sh));
    ws_bind(sh.ws, wsp, L);
#if POA_EMU_LANES == 32
    std::vector<char> ring((size_t)std::max<int>(P16_SMEM_BYTES, p16_mw_smem_bytes<8>()) + 16);
    sh.ring = ring.data();
    sh.ring_bytes = POA_EMU_NW == 1 ? P16_SMEM_BYTES : p16_mw_smem_bytes<POA_EMU_NW>();
#endif
#if POA_EMU_LANES > 1
    poa_emu::run_warp([&]() { poa_block<POA_EMU_NW>(sh, dp, B, L, O, 0, wsp); });
#else
    poa_block<1>(sh, dp, B, L, O, 0, wsp);
#endif
    if (getenv("POA_EMU_VERBOSE")) fprintf(stderr, "[emu] packed-16 alignments: %llu of %d\n", phase[PH_SPARE], n_seq > 0 ? n_seq - 1 : 0);
    if (hdr[H_STATUS] != ST_OK) { fprintf(stderr, "[emu] block status %d\n", hdr[H_STATUS]); *n_out = -hdr[H_STATUS]; return nullptr; }
    if (g_want_wire) {
        int32_t *out = (int32_t *)malloc(sizeof(int32_t) * (HDR_WORDS + used + 1));
        memcpy(out, hdr.data(), sizeof(int32_t) * HDR_WORDS);
        memcpy(out + HDR_WORDS, arena.data(), sizeof(int32_t) * used);
        *n_out = (int64_t)(HDR_WORDS + used);
        return out;
    }
    // wire format -> flat arrays (the decoder the C ABI uses) -> canonical dump
    const int n = hdr[H_N_NODE], ns = hdr[H_N_SEQ];
    const long long in_tot = hdr[H_IN_TOT], out_tot = hdr[H_OUT_TOT], aln_tot = hdr[H_ALN_TOT], path_tot = hdr[H_PATH_TOT], cig_tot = hdr[H_CIG_TOT];
    const int cons_len = hdr[H_CONS_LEN], msa_len = hdr[H_MSA_LEN], msa_rows = hdr[H_MSA_ROWS];
    const int *o = arena.data() + ((unsigned long long)(unsigned)hdr[H_OFF_LO] | ((unsigned long long)(unsigned)hdr[H_OFF_HI] << 32));
    DecodedBlock dec;
    if (!wire_header_ok(hdr.data()) || !wire_decode(hdr.data(), o, dec)) { fprintf(stderr, "[emu] corrupt wire body\n"); return nullptr; }
    const long long body = 4LL * n + 2 * in_tot + 2 * out_tot + aln_tot + ns + path_tot + (cons_len > 0 ? cons_len : 0);  // base .. cons_node
    pd_buf_t b = {0, 0, 0};
    for (int i = 0; i < PD_HEADER_LEN; ++i) pd_push(&b, 0);
    for (long long i = 0; i < body; ++i) pd_push(&b, dec.buf[(size_t)i]);
    for (long long i = 0; i < (long long)msa_rows * (msa_len > 0 ? msa_len : 0); ++i) pd_push(&b, dec.msa[i]);
    for (long long i = 0; i < 2LL * ns; ++i) pd_push(&b, dec.buf[(size_t)(dec.best + i)]);
    for (long long i = 0; i < 2 * cig_tot; ++i) pd_push(&b, dec.cig[i]);
    b.d[PD_MAGIC] = POA_DUMP_MAGIC; b.d[PD_N_NODE] = n; b.d[PD_N_SEQ] = ns;
    b.d[PD_CONS_LEN] = cons_len; b.d[PD_MSA_LEN] = msa_len; b.d[PD_MSA_ROWS] = msa_rows;
    b.d[PD_N_IN_TOT] = (int)in_tot; b.d[PD_N_OUT_TOT] = (int)out_tot; b.d[PD_N_ALN_TOT] = (int)aln_tot;
    b.d[PD_PATH_TOT] = (int)path_tot; b.d[PD_CIGAR_TOT] = (int)cig_tot;
    b.d[PD_INBAND_LO] = hdr[H_INBAND_LO]; b.d[PD_INBAND_HI] = hdr[H_INBAND_HI];
    b.d[PD_EDGE_ROWS_LO] = hdr[H_EDGE_LO]; b.d[PD_EDGE_ROWS_HI] = hdr[H_EDGE_HI];
    *n_out = b.n;
    return b.d;
}
