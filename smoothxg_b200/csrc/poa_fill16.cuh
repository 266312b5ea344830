// poa_fill16.cuh -- packed-int16 DP fill (included by poa_core.cuh inside namespace poa).
//
// Same recurrence as fill<NW, short>() (abPOA's convex row kernel, deps/abPOA/src/abpoa_align_simd.c:935-1074,
// first row :617-688, band :1107-1130), different machine mapping:
//
//  * two cells per 32-bit register, evaluated with Blackwell's packed 16-bit integer instructions
//    (VIADD.16x2, VIMNMX.S16x2, VIMNMX3.S16x2, VIADDMNMX.S16x2 -- __vadd2/__vmaxs2/__vimax3_s16x2/__viaddmax_s16x2);
//  * columns are processed in absolute 256-column chunks.  In a chunk, lane t holds columns
//    c0 + 4t .. c0 + 4t+3 in the low halves of four registers and c0 + 128 + 4t .. +3 in the high halves, so
//    the "previous column" operand of the match state is the neighbouring register (one shuffle per chunk
//    for register 0) and both halves run the horizontal-gap recurrences at once;
//  * F1/F2 are computed exactly: a 4-cell chain per lane, then a max-plus scan of the 64 lane-halves
//    (G = F + e*position turns the decaying recurrence into a running maximum), then a per-cell fix-up;
//  * a row is stored as [plane H,E1,E2][chunk][lane][4 words] = 512-byte chunk-planes, so every predecessor read and
//    every store is one fully coalesced 128-bit access per lane, and rows of different bands line up without
//    shifts because chunks are in absolute column coordinates.
//
// Cells of a stored chunk that lie outside the row's band [beg,end] hold inf_min in the H, E1 and E2 planes
// (what successors and the traceback may read); the F planes are not stored at all (see P16_PLANES).
#if POA_WARP == 32

#ifdef POA_HOST_EMU
static inline unsigned p_pack(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }
static inline int p_lo(unsigned x) { return (int)(short)(x & 0xffffu); }
static inline int p_hi(unsigned x) { return (int)(short)(x >> 16); }
static inline unsigned p_add(unsigned a, unsigned b) { return p_pack(p_lo(a) + p_lo(b), p_hi(a) + p_hi(b)); }
static inline unsigned p_max(unsigned a, unsigned b) { return p_pack(imax(p_lo(a), p_lo(b)), imax(p_hi(a), p_hi(b))); }
static inline unsigned p_min(unsigned a, unsigned b) { return p_pack(imin(p_lo(a), p_lo(b)), imin(p_hi(a), p_hi(b))); }
static inline unsigned p_max3(unsigned a, unsigned b, unsigned c) { return p_max(p_max(a, b), c); }
static inline unsigned p_minu(unsigned a, unsigned b) { unsigned lo = (a & 0xffffu) < (b & 0xffffu) ? (a & 0xffffu) : (b & 0xffffu), hi = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16); return lo | (hi << 16); }
static inline unsigned p_addmax(unsigned a, unsigned b, unsigned c) { return p_max(p_add(a, b), c); }
static inline unsigned p_signmask(unsigned d) { return (p_lo(d) < 0 ? 0xffffu : 0u) | (p_hi(d) < 0 ? 0xffff0000u : 0u); }
static inline unsigned p_swap(unsigned x) { return (x >> 16) | (x << 16); }
static inline int p_ctz(unsigned x) { return __builtin_ctz(x); }
static inline int p_clz(unsigned x) { return __builtin_clz(x); }
static inline unsigned p_lolo(unsigned a, unsigned b) { return (a & 0xffffu) | (b << 16); }  // (a.lo, b.lo)
#else
POA_D unsigned p_pack(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }
POA_D int p_lo(unsigned x) { return (int)(short)(x & 0xffffu); }
POA_D int p_hi(unsigned x) { return (int)x >> 16; }
POA_D unsigned p_add(unsigned a, unsigned b) { return __vadd2(a, b); }
POA_D unsigned p_max(unsigned a, unsigned b) { return __vmaxs2(a, b); }
POA_D unsigned p_min(unsigned a, unsigned b) { return __vmins2(a, b); }
POA_D unsigned p_max3(unsigned a, unsigned b, unsigned c) { return __vimax3_s16x2(a, b, c); }
POA_D unsigned p_minu(unsigned a, unsigned b) { return __vminu2(a, b); }
POA_D unsigned p_addmax(unsigned a, unsigned b, unsigned c) { return __viaddmax_s16x2(a, b, c); }
POA_D unsigned p_signmask(unsigned d) { return __byte_perm(d, 0u, 0xbb99u); }  // PRMT sign-replicate: 0xffff per negative half
POA_D unsigned p_swap(unsigned x) { return __byte_perm(x, 0u, 0x1032u); }
POA_D int p_ctz(unsigned x) { return __ffs((int)x) - 1; }
POA_D int p_clz(unsigned x) { return __clz((int)x); }
POA_D unsigned p_lolo(unsigned a, unsigned b) { return __byte_perm(a, b, 0x5410u); }
#endif

constexpr int P16_CW = 256;          // columns per chunk
constexpr int P16_CPB = 512;         // bytes per chunk-plane
#ifndef POA_P16_SMCH
#define POA_P16_SMCH 6
#endif
#ifndef POA_P16_QCH
#define POA_P16_QCH 4  // 16 resident blocks x (ring + profile stage + metadata + 1 KB reserved) must fit 228 KB of shared memory
#endif
// Stored planes per row: H, E1, E2 -- what successor rows and the traceback's M / E moves read.  The horizontal-gap planes
// F1 / F2 are NOT stored: no later row reads them, and the traceback needs them only on the rare insertion steps, where
// p16_row_f() below recomputes them for one row with the fill's own arithmetic (2 of 5 planes = 40 % of the DP's DRAM writes).
#ifdef POA_EXTRA_PLANE  // experiment only: one more (never read) plane per row, to see how far DRAM writes bound the fill
constexpr int P16_PLANES = 4;
#else
constexpr int P16_PLANES = 3;
#endif
constexpr int P16_SMCH = POA_P16_SMCH;  // chunks of the previous row kept in shared memory (H, E1, E2)
constexpr int P16_RING_BYTES = P16_SMCH * 3 * P16_CPB;
constexpr int P16_QCH = POA_P16_QCH;   // profile chunks of the row in flight staged in shared memory
constexpr int P16_QBUF_OFF = P16_RING_BYTES, P16_META_OFF = P16_QBUF_OFF + P16_QCH * P16_CPB;
constexpr int P16_SLOT = 144;                                  // one row's prefetched metadata: nine 16-byte cells (layout: fill_p16)
constexpr int P16_MW_SLOT = 160;                               // the multi-warp fill also prefetches the third / fourth predecessor
constexpr int P16_OUTS_OFF = P16_META_OFF + 2 * P16_SLOT;     // + 128 B: the row's successor rows
constexpr int P16_LANEC_OFF = P16_OUTS_OFF + 128;          // + 512 B: four lane-dependent packed constants per lane (fill_p16)
constexpr int P16_SMEM_BYTES = P16_LANEC_OFF + 512;

// address of cell (plane, j) of a row stored in the chunked layout; pm = {first chunk-plane of the row, beg, end, _}
POA_D const short *cell_ptr16(const Ws &w, const int4 &pm, int plane, int j) {
    const int cb = pm.y >> 8, nch = (pm.z >> 8) - cb + 1;
    const int u = j & 255;
    return reinterpret_cast<const short *>(w.slab) + ((long long)pm.x + (long long)plane * nch + ((j >> 8) - cb)) * 256 + ((u & 127) << 1) + (u >> 7);
}

// the ring lives in shared memory: explicit ld/st.shared on the 32-bit shared-space address (a generic pointer
// kept in the Shared struct would compile to slower generic LD/ST)
#ifdef POA_HOST_EMU
typedef char *ring_ptr_t;
static inline ring_ptr_t ring_base(char *p, int lane) { return p + lane * 16; }
static inline uint4 ring_ld(ring_ptr_t p, unsigned off) { return *reinterpret_cast<const uint4 *>(p + off); }
static inline void ring_st(ring_ptr_t p, unsigned off, unsigned a, unsigned b, unsigned c, unsigned d) { uint4 u; u.x = a; u.y = b; u.z = c; u.w = d; *reinterpret_cast<uint4 *>(p + off) = u; }
#else
typedef unsigned ring_ptr_t;
POA_D ring_ptr_t ring_base(char *p, int lane) { return (unsigned)__cvta_generic_to_shared(p) + lane * 16; }
POA_D uint4 ring_ld(ring_ptr_t p, unsigned off) {
    uint4 u;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(p + off));
    return u;
}
POA_D void ring_st(ring_ptr_t p, unsigned off, unsigned a, unsigned b, unsigned c, unsigned d) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(p + off), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
#endif

// Asynchronous global -> shared copies (cp.async / LDGSTS).  Everything the next row or the next chunk needs
// from global memory is staged this way: completion is tracked by cp.async groups, not by the load scoreboards,
// so an in-flight prefetch can never stall an unrelated instruction that happens to share a scoreboard slot.
// `ring_ptr_t` arithmetic: shared-space byte address (device) / host pointer (emulation).
#ifdef POA_HOST_EMU
static inline void cpa16(ring_ptr_t s, unsigned off, const void *g, bool l2_only) { (void)l2_only; memcpy(s + off, g, 16); }
static inline void cpa16z(ring_ptr_t s, unsigned off, const void *g, unsigned n) { if (n) memcpy(s + off, g, 16); else memset(s + off, 0, 16); }
static inline void cpa4(ring_ptr_t s, unsigned off, const void *g) { memcpy(s + off, g, 4); }
static inline void cpa_commit() {}
static inline void cpa_wait_pending(int n) { (void)n; }
static inline int ring_ld32(ring_ptr_t p, unsigned off) { int v; memcpy(&v, p + off, 4); return v; }
#else
POA_D void cpa16(ring_ptr_t s, unsigned off, const void *g, bool l2_only) {
    if (l2_only) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s + off), "l"(g) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(s + off), "l"(g) : "memory");
}
// 16-byte copy at L2 that moves `n` = 16 or 0 source bytes and zero-fills the rest: lets every lane of a gather issue the same
// instruction whether or not its item exists
POA_D void cpa16z(ring_ptr_t s, unsigned off, const void *g, unsigned n) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(s + off), "l"(g), "r"(n) : "memory"); }
POA_D void cpa4(ring_ptr_t s, unsigned off, const void *g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(s + off), "l"(g) : "memory"); }
POA_D void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
POA_D void cpa_wait_pending(int n) {  // wait until at most n of the most recent groups are still in flight (uniform n)
    switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    }
}
POA_D int ring_ld32(ring_ptr_t p, unsigned off) { int v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(p + off)); return v; }
#endif

POA_D uint4 p16_ld(const char *p) { return *reinterpret_cast<const uint4 *>(p); }
// band inputs are updated with L2 reductions (poa_red_max/min), so they are read at L2, never from a stale L1 line
#ifdef POA_HOST_EMU
static inline int p16_ldcg(const int *p) { return *p; }
#else
POA_D int p16_ldcg(const int *p) { return __ldcg(p); }
#endif
POA_D void p16_st(char *p, unsigned a, unsigned b, unsigned c, unsigned d) {
    uint4 u; u.x = a; u.y = b; u.z = c; u.w = d;
#if defined(POA_ST_CS) && !defined(POA_HOST_EMU)
    __stcs(reinterpret_cast<uint4 *>(p), u);  // row planes are written once and (the ring aside) read much later or never: evict first
#else
    *reinterpret_cast<uint4 *>(p) = u;
#endif
}

// A value the compiler may not move or speculate: keeps the "-inf" fill of an out-of-range predecessor chunk on its (rare) branch
// instead of twelve unconditional register initialisations in front of every chunk's loads.
#ifdef POA_HOST_EMU
static inline unsigned p16_pinned(unsigned v) { return v; }
static inline const char *p16_pinned_ptr(const char *p) { return p; }
#else
POA_D unsigned p16_pinned(unsigned v) { unsigned r; asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(v)); return r; }
POA_D const char *p16_pinned_ptr(const char *p) { asm volatile("" : "+l"(p)); return p; }  // the compiler may not re-derive it per use
#endif

// Can this alignment run in packed 16-bit arithmetic without any intermediate leaving the int16 range?
// (scores <= qlen*match; the scan adds at most e*256 of position offset on top.)
POA_D bool p16_eligible(const DevParams &P, int qlen) {
    const int emax = imax(P.e1, P.e2);
    return P.p16_ok && (long long)qlen * P.match + (long long)emax * (P16_CW + 8) + 64 <= 32767;
}

// Compile-time chunk count of one pass of the row loop (fill_p16): generic lambdas take it as a value of this type.
template <int N> struct p16_n { static constexpr int value = N; };
template <bool B> struct p16_b { static constexpr bool value = B; };
#ifndef POA_P16_ILP
#define POA_P16_ILP 1  // chunks of a row evaluated side by side by one warp; 1 is what ships (see below), 2..4 are experiments
#endif

// fill_p16: one warp, rows in index order, the chunks of a row taken POA_P16_ILP at a time.
//
// The row loop is written for N chunks per pass (`pass`, a generic lambda over a compile-time N): a row is one dependent chain
// per chunk -- predecessor loads, the shifted match operand, a 4-cell recurrence per lane, a five-round shuffle scan, the
// fix-up -- and consecutive rows depend on each other through the adaptive band (a row's band needs the arg-max columns of
// the whole previous row), so a warp that walks one chunk at a time has little independent work to issue while a shuffle or
// a load is in flight (52 % of issue slots used).  The chunks of ONE row are independent up to the scalar F carry that enters
// a chunk from its left neighbour, and that carry is applied AFTER the lane scan, so N chunks can run as interleaved
// straight-line code.  MEASURED on configs[2] (profiles/r02_ilp_variants.json): N = 2 cuts long-scoreboard, short-scoreboard
// and fixed-latency stalls per issued instruction from 2.22 / 0.60 / 1.79 to 1.40 / 0.36 / 1.53, but the pass body no longer
// fits the SM's instruction caches (726 instead of 373 instructions; `no_instruction` stalls rise from 0.49 to 1.41 per issue)
// and it needs 168 registers (12 resident warps per SM instead of 16): 188-211 Gcells/s against 250 for N = 1; N = 3 and 4 are
// slower still (150 / 122).  So N = 1 ships; the switch stays for the next machine with a larger L0/L1.5 I-cache.
// PRESET: the scoring parameters are smoothxg's defaults (match 1, mismatch 4, gaps 6,2,26,1; DevParams::p16_default): the packed
// constants below are then immediates of the instructions that use them instead of nine constant-bank loads per chunk pass.
template <int NW, bool LOCAL, bool PRESET>
POA_D void fill_p16(Shared &sh, const DevParams &P, const WsLayout &L, char *const wsb, const uint8_t *q, int qlen, const int pn) {
    const int lane = poa_tid();
    const int n_node = sh.n_node;
    const int rows = n_node - 1;  // the sink row is never filled
    const int inf_min = PRESET ? -31717 : inf_min_of<short>(P);
    constexpr bool local = LOCAL;
    const int wb = local ? -1 : P.wb;  // abpoa_align.c:158
#ifdef POA_HOST_EMU
    const int bw = wb < 0 ? qlen : wb + (int)(P.wf * qlen);  // abpoa_align_simd.c:474
#else
    const int bw = wb < 0 ? qlen : wb + (int)__fmul_rn(P.wf, (float)qlen);
#endif
    const int e1 = PRESET ? 2 : P.e1, e2 = PRESET ? 1 : P.e2, oe1 = PRESET ? 8 : P.oe1, oe2 = PRESET ? 27 : P.oe2;
    // Workspace pointers are formed from the CTA's workspace base (a kernel parameter plus blockIdx.x * stride: warp-uniform) and
    // the layout's offsets (constant bank), NOT read back from the Ws struct in shared memory: values loaded from memory are
    // not provably uniform, so a dozen 64-bit pointers would each pin two vector registers for the whole alignment -- with
    // 128 registers per thread that was what pushed loop invariants onto the stack (two local-memory reloads per row, 9 % of
    // all stall samples in the ncu source view).  Uniform values live in uniform registers or are rematerialised for free.
    char *const slab = wsb + L.o_slab, *const qp = wsb + L.o_qp;
    const int4 *const rowinfo = reinterpret_cast<const int4 *>(wsb + L.o_rowinfo);
    int4 *const rowmeta = reinterpret_cast<int4 *>(wsb + L.o_rowmeta);
    const int *const pool_row = reinterpret_cast<const int *>(wsb + L.o_pool_row), *const rr = reinterpret_cast<const int *>(wsb + L.o_rr);
    int *const mplr = reinterpret_cast<int *>(wsb + L.o_mplr), *const mprr = reinterpret_cast<int *>(wsb + L.o_mprr);
    const uint8_t *const rbase = reinterpret_cast<const uint8_t *>(wsb + L.o_rbase);
    const int *const fp = reinterpret_cast<const int *>(wsb + L.o_tmp0);  // build_rows(): row of the first predecessor
    const int *const sp = reinterpret_cast<const int *>(wsb + L.o_tmp1);  // build_rows(): row of the second predecessor, -1 if there is none
#ifndef POA_HOST_EMU
    __builtin_assume(__isGlobal(slab)); __builtin_assume(__isGlobal(qp)); __builtin_assume(__isGlobal(rowinfo));
    __builtin_assume(__isGlobal(rowmeta)); __builtin_assume(__isGlobal(pool_row)); __builtin_assume(__isGlobal(rr));
    __builtin_assume(__isGlobal(mplr)); __builtin_assume(__isGlobal(mprr)); __builtin_assume(__isGlobal(rbase));
    __builtin_assume(__isGlobal(q)); __builtin_assume(__isGlobal(fp)); __builtin_assume(__isGlobal(sp));
#endif
    // chunk-planes the slab holds / already used: 32-bit (a row's offset is stored as one 32-bit chunk-plane index anyway)
    const unsigned slab_units = L.slab_planes;
    unsigned used = 0;
    long long inband = 0, edge_rows = 0;

    // packed constants
    const int emax = imax(e1, e2);
    const int negl = -32768 + 8 * emax + 8;  // below every value a band cell can take, never wraps when used
    (void)negl;
    // uniform ones come from the constant bank (DevParams::pk_*, build_params()); the lane-dependent four stay in registers
#define P16_PK(v) ((((unsigned)(v)) & 0xffffu) | (((unsigned)(v)) << 16))
#define INFP (PRESET ? P16_PK(-31717) : P.pk_inf)
#define NEGLP (PRESET ? P16_PK(-32744) : P.pk_negl)
#define NOE1 (PRESET ? P16_PK(-8) : P.pk_noe1)
#define NOE2 (PRESET ? P16_PK(-27) : P.pk_noe2)
#define NE1 (PRESET ? P16_PK(-2) : P.pk_ne1)
#define NE2 (PRESET ? P16_PK(-1) : P.pk_ne2)
#define NE1_2 (PRESET ? P16_PK(-4) : P.pk_ne1_2)
#define NE1_3 (PRESET ? P16_PK(-6) : P.pk_ne1_3)
#define NE2_2 (PRESET ? P16_PK(-2) : P.pk_ne2_2)
#define NE2_3 (PRESET ? P16_PK(-3) : P.pk_ne2_3)
#define NCW1 (PRESET ? P16_PK(-512) : P.pk_ncw1)
#define NCW2 (PRESET ? P16_PK(-256) : P.pk_ncw2)
    const unsigned ZERO = 0u;
    const unsigned OFF1 = p_pack(e1 * 4 * (lane + 1), e1 * 4 * (lane + 33)), NOFF1 = p_pack(-e1 * 4 * lane, -e1 * 4 * (lane + 32));
    const unsigned OFF2 = p_pack(e2 * 4 * (lane + 1), e2 * 4 * (lane + 33)), NOFF2 = p_pack(-e2 * 4 * lane, -e2 * 4 * (lane + 32));
    const int f0_1 = imax(inf_min - oe1, inf_min - e1), f0_2 = imax(inf_min - oe2, inf_min - e2);
    // The four lane-dependent constants of the scan live in shared memory (one 16-byte slice per lane) and are re-read by every
    // chunk pass with ONE ld.shared.v4: with 128 registers per thread they do not stay in registers, and the compiler otherwise
    // rebuilds them from the lane number in every pass (22 of the pass's 243 instructions).
    ring_st(ring_base(sh.ring, lane), P16_LANEC_OFF, OFF1, NOFF1, OFF2, NOFF2);

    // ---- query profile in the chunked layout: qp[base][chunk][lane] (abpoa_align_simd.c:531-546)
    const int nchq = (qlen >> 8) + 1;
    for (int b = 0; b < 5; ++b)
        for (int c = 0; c < nchq; ++c) {
            unsigned v[4];
            for (int r = 0; r < 4; ++r) {
                const int jl = c * P16_CW + lane * 4 + r, jh = jl + 128;
                const int sl = (jl == 0 || jl > qlen) ? 0 : P.mat[b * 5 + q[jl - 1]];
                const int s2 = jh > qlen ? 0 : P.mat[b * 5 + q[jh - 1]];
                v[r] = p_pack(sl, s2);
            }
            p16_st(qp + ((long long)(b * nchq + c)) * P16_CPB + lane * 16, v[0], v[1], v[2], v[3]);
        }

    // ---- row 0 (abpoa_align_simd.c:617-688)
    {
        int end0;
        if (wb >= 0) {
            if (lane == 0) {
                mplr[0] = 0; mprr[0] = 0;
                const int4 ri = rowinfo[0];
                for (int k = 0; k < ri.w; ++k) { int o = pool_row[ri.z + k]; mplr[o] = 1; mprr[o] = 1; }
            }
            end0 = imin(qlen, imax(0, rr[0]) + bw);
        } else end0 = qlen;
        const int nch = (end0 >> 8) + 1;
        if ((unsigned)(P16_PLANES * nch) > slab_units) { if (lane == 0) sh.err = ST_ESLAB; sync_block<NW>(); return; }
        if (lane == 0) rowmeta[0] = poa_make_int4(0, 0, end0, 0);
        for (int c = 0; c < nch; ++c) {
            unsigned v[5][4];
            for (int r = 0; r < 4; ++r) {
                int x[2][5];
                for (int hf = 0; hf < 2; ++hf) {
                    const int j = c * P16_CW + hf * 128 + lane * 4 + r;
                    int *y = x[hf];
                    if (j > end0) { y[0] = y[1] = y[2] = y[3] = y[4] = inf_min; }
                    else if (local) { y[0] = y[1] = y[2] = y[3] = y[4] = 0; }
                    else if (j == 0) { y[0] = 0; y[1] = -oe1; y[2] = -oe2; y[3] = y[4] = inf_min; }
                    else { y[3] = -P.o1 - e1 * j; y[4] = -P.o2 - e2 * j; y[0] = imax((int)(short)y[3], (int)(short)y[4]); y[1] = y[2] = inf_min; }
                }
                for (int p = 0; p < 5; ++p) v[p][r] = p_pack(x[0][p], x[1][p]);
            }
            for (int p = 0; p < 3; ++p)  // H, E1, E2 (row 0 is never a traceback row: the walk stops at i == 0)
                p16_st(slab + ((long long)p * nch + c) * P16_CPB + lane * 16, v[p][0], v[p][1], v[p][2], v[p][3]);
        }
        used = (unsigned)(P16_PLANES * nch);
        inband += end0 + 1;
        sync_block<NW>();
    }
    int best_score = inf_min, best_i = 0, best_j = 0;
    const int pshift = 31 - p_clz((unsigned)pn);  // pn is the reference build's lane count, a power of two
    char *const slab_lane = slab + lane * 16;
    const bool track = local || wb >= 0;
    int4 prev_meta = rowmeta[0];  // {chunk-plane index, beg, end} of the row evaluated last
    int prev_left = 0, prev_right = 0;  // its arg-max columns (band propagation, forwarded in registers)
    // The previous row's H/E1/E2 chunks also stay in shared memory (each lane's own 16-byte slices; slot = chunk
    // % P16_SMCH), so the common "predecessor = row just evaluated" read never leaves the SM.
    const ring_ptr_t ring = ring_base(sh.ring, lane);
    bool prev_res = false;  // is the previous row in the ring? (row 0 is not; rows wider than the ring are not)
    // Metadata of the next row is gathered into shared memory one row ahead by cp.async, one item per lane (the numbers
    // of its first and second predecessor two ahead, so that those rows' descriptors can be fetched one ahead too).
    // A slot is nine 16-byte cells, each the aligned 16-byte window of its array that holds the item (read at L2: the band
    // inputs are updated by reductions): +0 rowinfo, +16 rowmeta[first pred], +32 bases (16 rows), +48 mplr, +64 mprr (4 rows),
    // +80 rowmeta[second pred], +96 fp (4 rows; the next row's entry is used), +112 rr (4 rows), +128 sp (4 rows, next row's).
    const ring_ptr_t sm = ring_base(sh.ring, 0);
    // each of lanes 0..8 owns one cell: its array, element size and slot offset are fixed for the whole alignment
    const char *gsrc = nullptr; unsigned gdst = 0; int gshift = 0, gmask = ~0;  // element index -> byte offset of its 16-byte window
    if (lane == 0) { gsrc = (const char *)rowinfo; gdst = 0; gshift = 4; }
    else if (lane == 1) { gsrc = (const char *)rowmeta; gdst = 16; gshift = 4; }
    else if (lane == 2) { gsrc = (const char *)rbase; gdst = 32; gshift = 0; gmask = ~15; }
    else if (lane == 3) { gsrc = (const char *)fp; gdst = 96; gshift = 2; gmask = ~3; }
    else if (lane == 4 && wb >= 0) { gsrc = (const char *)rr; gdst = 112; gshift = 2; gmask = ~3; }
    else if (lane == 5 && wb >= 0) { gsrc = (const char *)mplr; gdst = 48; gshift = 2; gmask = ~3; }
    else if (lane == 6 && wb >= 0) { gsrc = (const char *)mprr; gdst = 64; gshift = 2; gmask = ~3; }
    else if (lane == 7) { gsrc = (const char *)rowmeta; gdst = 80; gshift = 4; }
    else if (lane == 8) { gsrc = (const char *)sp; gdst = 128; gshift = 2; gmask = ~3; }
    // which row the lane's item belongs to (0: row n1, 1 / 2: its first / second predecessor, 3: the row after n1): fixed per lane
    // too, so the gather is ONE copy instruction for the nine lanes, no branches (an item that does not exist is zero-filled)
    const int gsel = lane == 1 ? 1 : lane == 7 ? 2 : (lane == 3 || lane == 8) ? 3 : 0;
    auto gather = [&](const int n1, const int np0_n1, const int nsp_n1, const int cur) {  // row n1 into its slot; `cur`: row being evaluated
        if (n1 >= rows || gsrc == nullptr) return;
        const int idx = gsel == 1 ? np0_n1 : gsel == 2 ? nsp_n1 : n1 + (gsel == 3 ? 1 : 0);
        // a predecessor's descriptor is fetched only if that row is complete (rows >= cur: forwarded in registers instead)
        const bool live = (unsigned)idx < (unsigned)((gsel == 1 || gsel == 2) ? cur : rows);
        cpa16z(sm, P16_META_OFF + (unsigned)(n1 & 1) * P16_SLOT + gdst, gsrc + (live ? (unsigned)(idx & gmask) << gshift : 0u), live ? 16u : 0u);
    };
    int np0 = fp[rows > 1 ? 1 : 0];   // first predecessor of the row about to be evaluated
    int nsp = rows > 1 ? sp[1] : -1;  // its second predecessor, or -1
    gather(1, np0, nsp, 1);
    cpa_commit();

    // ---- rows in index order (abpoa_align_simd.c:1205-1221)
    for (int i = 1; i < rows; ++i) {
        cpa_wait_pending(0);
        sync_block<NW>();  // the slot was filled by other lanes' copies
        const unsigned slot = P16_META_OFF + (unsigned)(i & 1) * P16_SLOT;
        const uint4 ri_u = ring_ld(sm, slot), npm_u = ring_ld(sm, slot + 16), spm_u = ring_ld(sm, slot + 80);
        const int4 ri = poa_make_int4((int)ri_u.x, (int)ri_u.y, (int)ri_u.z, (int)ri_u.w);  // {in_off, in_n, out_off, out_n}
        const unsigned w0 = slot + 4u * (unsigned)(i & 3), w1 = slot + 4u * (unsigned)((i + 1) & 3);  // this / the next row's word of a window
        const int rb = (ring_ld32(sm, slot + 32 + 4u * (unsigned)((i >> 2) & 3)) >> (8 * (i & 3))) & 0xff, p0 = np0, s0 = nsp;
        const int nnp0 = i + 1 < rows ? ring_ld32(sm, w1 + 96) : 0;
        const int nnsp = i + 1 < rows ? ring_ld32(sm, w1 + 128) : -1;
        const int r = ring_ld32(sm, w0 + 112);
        int ml = ring_ld32(sm, w0 + 48), mr = ring_ld32(sm, w0 + 64);
        // the first two predecessors' row descriptors: from registers when it is the row just evaluated (the common case)
        const int4 pm0 = p0 == i - 1 ? prev_meta : poa_make_int4((int)npm_u.x, (int)npm_u.y, (int)npm_u.z, (int)npm_u.w);
        const int4 pm1 = s0 == i - 1 ? prev_meta : poa_make_int4((int)spm_u.x, (int)spm_u.y, (int)spm_u.z, (int)spm_u.w);
        // profile chunks of this row: staged now for the chunk range the previous row covered plus one (bands move
        // slowly), so the copies overlap the band computation below; corrected after it if the guess was wrong
        int scb = prev_meta.y >> 8, sce = imin(imin((prev_meta.z >> 8) + 1, scb + P16_QCH - 1), nchq - 1);
        {
            // at most P16_QCH copies: written as predicated straight-line code (a counted loop makes the compiler emit unroll-by-16 /
            // 8 / 4 bodies for a trip count it cannot see is <= 4: 60 instructions of row-loop code that never run)
            const char *qg = p16_pinned_ptr(qp + (size_t)(unsigned)(rb * nchq + scb) * P16_CPB + lane * 16);  // one address, four immediate offsets
            const int nq = sce - scb;
#pragma unroll
            for (int k = 0; k < P16_QCH; ++k) if (k <= nq) cpa16(sm, P16_QBUF_OFF + (unsigned)k * P16_CPB + lane * 16, qg + (size_t)k * P16_CPB, false);
            cpa_commit();
        }
        // The profile copies are committed FIRST and the gather of the next row's metadata after them: the first chunk pass then
        // waits for "all but the newest group" = the profile only, and the gather has the whole row to land.
        gather(i + 1, nnp0, nnsp, i);
        // rows this row hands its arg-max columns to (one per lane; staged now, used after the last chunk)
        if (lane < ri.w) cpa4(sm, P16_OUTS_OFF + lane * 4, &pool_row[ri.z + lane]);
        cpa_commit();
        np0 = nnp0; nsp = nnsp;
        int q_newer = 1;  // cp.async groups committed after the profile's
        int beg, end;
        if (wb < 0) { beg = 0; end = qlen; }
        else {  // abpoa_align.h:34-35, abpoa_align_simd.c:946-960
            int min_pre_beg = pm0.y;
            bool from_prev = p0 == i - 1;
            if (ri.y > 1) { from_prev |= s0 == i - 1; min_pre_beg = imin(min_pre_beg, pm1.y); }
#pragma unroll 1
            for (int k = 2; k < ri.y; ++k) {
                const int pk = pool_row[ri.x + k];
                from_prev |= pk == i - 1;
                min_pre_beg = imin(min_pre_beg, rowmeta[pk].y);
            }
            if (from_prev) { ml = imin(ml, prev_left + 1); mr = imax(mr, prev_right + 1); }  // abpoa_align_simd.c:1121-1130
            beg = imax(0, imin(ml, r) - bw);
            end = imin(qlen, imax(mr, r) + bw);
            if ((beg >> pshift) < (min_pre_beg >> pshift)) beg = min_pre_beg;
        }
        if (end < beg) end = beg;
        const int cb = beg >> 8, ce = end >> 8, nch = ce - cb + 1;
        if (used + (unsigned)(P16_PLANES * nch) > slab_units) {
            if (lane == 0) sh.err = ST_ESLAB;
            sync_block<NW>();
            return;
        }
        const unsigned roff = used;
        used += (unsigned)(P16_PLANES * nch);
        inband += end - beg + 1;
        edge_rows += (long long)ri.y * (end - beg + 1);

        // F entering column cb*256 such that F[beg] comes out as f0 (cells left of beg are masked to inf_min)
        unsigned carry1 = p_pack(f0_1 + e1 * (beg - cb * P16_CW), f0_1 + e1 * (beg - cb * P16_CW));
        unsigned carry2 = p_pack(f0_2 + e2 * (beg - cb * P16_CW), f0_2 + e2 * (beg - cb * P16_CW));
        int rmx = INT_MIN, fc = cb, lc = cb;  // row maximum, first / last chunk attaining it
        const char *qrow = qp + (size_t)(unsigned)(rb * nchq) * P16_CPB + lane * 16;
        const bool cur_res = nch <= P16_SMCH;

        // the row's profile chunks are staged in shared memory
        bool qst = cb >= scb && ce <= sce;
        if (!qst && nch <= P16_QCH) {  // wrong guess (rare): let the speculative copy land, then stage the exact range over it
            cpa_wait_pending(0);
            {
                const char *qg = p16_pinned_ptr(qrow + (size_t)(unsigned)cb * P16_CPB);
#pragma unroll
                for (int k = 0; k < P16_QCH; ++k) if (k < nch) cpa16(sm, P16_QBUF_OFF + (unsigned)k * P16_CPB + lane * 16, qg + (size_t)k * P16_CPB, false);
            }
            cpa_commit();
            q_newer = 0;
            scb = cb; qst = true;
        }
        char *dst = slab_lane + (size_t)roff * P16_CPB;
        const size_t pstride = (size_t)(unsigned)nch * P16_CPB;
        // first predecessor (always present; for most rows the only one, and the row just evaluated)
        const int pcb0 = pm0.y >> 8, pce0 = pm0.z >> 8;
        const unsigned pn0 = (unsigned)(pce0 - pcb0 + 1);
        const bool ring0 = prev_res && p0 == i - 1;
        unsigned rslot = (unsigned)(cb % P16_SMCH) * (3 * P16_CPB);  // ring slot of chunk c: (c % P16_SMCH) chunk-plane triples in
        // H of the first predecessor at the column just left of the chunk about to be evaluated (lane 0's match operand);
        // after a chunk it is simply the last cell of the predecessor chunk just read (inf_min if that was out of range)
        int plast = inf_min;
        if (cb > pcb0 && cb <= pce0 + 1) plast = *reinterpret_cast<const short *>(slab + (size_t)((unsigned)pm0.x + (unsigned)(cb - pcb0)) * P16_CPB - 2);
        // the profile chunks were requested before the band computation: wait for them here, once, instead of testing "first pass?"
        // in every chunk pass (the gather committed after them stays in flight)
        if (qst) { if (q_newer) cpa_wait_pending(1); else cpa_wait_pending(0); }

        const bool bnd_first = (beg & (P16_CW - 1)) != 0, bnd_last = (end & (P16_CW - 1)) != P16_CW - 1;  // cells outside the band in the first / last chunk?
        // STAGED: the row's profile chunks are in shared memory (every row of a batch whose bands fit P16_QCH chunks); rows wider than
        // that read the profile from global memory in a second instantiation of the pass, which costs the common one nothing
        auto pass = [&](auto nc, auto staged, const int c) {  // chunks c .. c + N - 1 of row i, side by side
            constexpr int N = decltype(nc)::value;
            constexpr bool STAGED = decltype(staged)::value;
            unsigned M[N][4], A[N][4], B[N][4], H[N][4];
            // ---- first predecessor: its chunk (H shifted one column right for M, E1, E2 as they are), or inf_min
#pragma unroll
            for (int u = 0; u < N; ++u) {
                const int cu = c + u;
                uint4 h, a, b;
                if (cu >= pcb0 && cu <= pce0) {
                    if (ring0) {
                        const unsigned rs = rslot + (unsigned)u * (3 * P16_CPB), rsw = rs >= P16_SMCH * (3 * P16_CPB) ? rs - P16_SMCH * (3 * P16_CPB) : rs;
                        h = ring_ld(ring, rsw); a = ring_ld(ring, rsw + P16_CPB); b = ring_ld(ring, rsw + 2 * P16_CPB);
                    } else {
                        const unsigned idx = (unsigned)pm0.x + (unsigned)(cu - pcb0);
                        h = p16_ld(slab_lane + (size_t)idx * P16_CPB);
                        a = p16_ld(slab_lane + (size_t)(idx + pn0) * P16_CPB);
                        b = p16_ld(slab_lane + (size_t)(idx + 2 * pn0) * P16_CPB);
                    }
                } else {  // the chunk lies outside the predecessor's band (rare)
                    const unsigned inf = p16_pinned(INFP);
                    h.x = h.y = h.z = h.w = inf; a = h; b = h;
                }
                const unsigned rot = (unsigned)poa_shfl((int)h.w, (lane + 31) & 31);
                M[u][0] = lane == 0 ? p_pack(plast, p_lo(rot)) : rot; M[u][1] = h.x; M[u][2] = h.y; M[u][3] = h.z;
                plast = p_hi(rot);  // lane 0: the predecessor's last cell of this chunk (inf_min when the chunk was out of its range)
                A[u][0] = a.x; A[u][1] = a.y; A[u][2] = a.z; A[u][3] = a.w;
                B[u][0] = b.x; B[u][1] = b.y; B[u][2] = b.z; B[u][3] = b.w;
            }
            // ---- further predecessors in in_id order (abpoa_align_simd.c:966-1029)
#pragma unroll 1
            for (int k = 1; k < ri.y; ++k) {
                int pk = s0;
                int4 pm = pm1;
                if (k > 1) { pk = pool_row[ri.x + k]; pm = rowmeta[pk]; }
                const int pcb = pm.y >> 8, pce = pm.z >> 8;
                if (c + N - 1 < pcb || c > pce + 1) continue;
                const unsigned pn_ = (unsigned)(pce - pcb + 1);
                const bool in_ring = prev_res && pk == i - 1;
                int pl = inf_min;  // this predecessor's H at the column left of chunk c
                if (c > pcb && c <= pce + 1) pl = *reinterpret_cast<const short *>(slab + (size_t)((unsigned)pm.x + (unsigned)(c - pcb)) * P16_CPB - 2);
#pragma unroll
                for (int u = 0; u < N; ++u) {
                    const int cu = c + u;
                    uint4 h, a, b;
                    if (cu >= pcb && cu <= pce) {
                        if (in_ring) {
                            const unsigned rs = rslot + (unsigned)u * (3 * P16_CPB), rsw = rs >= P16_SMCH * (3 * P16_CPB) ? rs - P16_SMCH * (3 * P16_CPB) : rs;
                            h = ring_ld(ring, rsw); a = ring_ld(ring, rsw + P16_CPB); b = ring_ld(ring, rsw + 2 * P16_CPB);
                        } else {
                            const unsigned idx = (unsigned)pm.x + (unsigned)(cu - pcb);
                            h = p16_ld(slab_lane + (size_t)idx * P16_CPB);
                            a = p16_ld(slab_lane + (size_t)(idx + pn_) * P16_CPB);
                            b = p16_ld(slab_lane + (size_t)(idx + 2 * pn_) * P16_CPB);
                        }
                    } else {
                        const unsigned inf = p16_pinned(INFP);
                        h.x = h.y = h.z = h.w = inf; a = h; b = h;
                    }
                    const unsigned rot = (unsigned)poa_shfl((int)h.w, (lane + 31) & 31);
                    const unsigned s0_ = lane == 0 ? p_pack(pl, p_lo(rot)) : rot;
                    pl = p_hi(rot);
                    M[u][0] = p_max(M[u][0], s0_); M[u][1] = p_max(M[u][1], h.x); M[u][2] = p_max(M[u][2], h.y); M[u][3] = p_max(M[u][3], h.z);
                    A[u][0] = p_max(A[u][0], a.x); A[u][1] = p_max(A[u][1], a.y); A[u][2] = p_max(A[u][2], a.z); A[u][3] = p_max(A[u][3], a.w);
                    B[u][0] = p_max(B[u][0], b.x); B[u][1] = p_max(B[u][1], b.y); B[u][2] = p_max(B[u][2], b.z); B[u][3] = p_max(B[u][3], b.w);
                }
            }
            if (local && c == 0 && lane == 0) M[0][0] = p_max(M[0][0], p_pack(0, inf_min));  // abpoa_align_simd.c:974 (`first` = 0)
            // ---- H~ = max(M + profile, E1, E2) (abpoa_align_simd.c:1032-1050); the profile chunks are read as late as possible
#pragma unroll
            for (int u = 0; u < N; ++u) {
                uint4 qv;
                if (STAGED) qv = ring_ld(sm, P16_QBUF_OFF + (unsigned)(c + u - scb) * P16_CPB + lane * 16);
                else qv = p16_ld(qrow + (size_t)(unsigned)(c + u) * P16_CPB);
                H[u][0] = p_max3(p_add(M[u][0], qv.x), A[u][0], B[u][0]); H[u][1] = p_max3(p_add(M[u][1], qv.y), A[u][1], B[u][1]);
                H[u][2] = p_max3(p_add(M[u][2], qv.z), A[u][2], B[u][2]); H[u][3] = p_max3(p_add(M[u][3], qv.w), A[u][3], B[u][3]);
            }
            // ---- cells outside [beg,end]: only the row's first and last chunk can hold any
            const bool bnd = (c == cb && bnd_first) || (c + N - 1 == ce && bnd_last);
            unsigned m[N][4];
#pragma unroll
            for (int u = 0; u < N; ++u) { m[u][0] = 0; m[u][1] = 0; m[u][2] = 0; m[u][3] = 0; }
            if (bnd) {
#pragma unroll
                for (int u = 0; u < N; ++u) {
                    if (u != 0 && u != N - 1) continue;  // an inner chunk of the pass is an inner chunk of the row
                    const int c0 = (c + u) * P16_CW;
                    const int brel = imax(beg - c0, 0), erel = imin(end - c0, P16_CW - 1);
                    const unsigned da = p_pack(lane * 4 - brel, 128 + lane * 4 - brel), db = p_pack(erel - lane * 4, erel - 128 - lane * 4);
                    m[u][0] = p_signmask(p_min(da, db));
                    m[u][1] = p_signmask(p_min(p_add(da, 0x00010001u), p_add(db, 0xffffffffu)));
                    m[u][2] = p_signmask(p_min(p_add(da, 0x00020002u), p_add(db, 0xfffefffeu)));
                    m[u][3] = p_signmask(p_min(p_add(da, 0x00030003u), p_add(db, 0xfffdfffdu)));
#pragma unroll
                    for (int x = 0; x < 4; ++x) H[u][x] = (H[u][x] & ~m[u][x]) | (INFP & m[u][x]);
                }
            }
            // ---- horizontal gaps (abpoa_align_simd.c:1052-1059): per-lane chains and the lane-half scans of all N chunks
            unsigned l[N][3], k_[N][3], g1[N], g2[N];
            const uint4 lc4 = ring_ld(ring, P16_LANEC_OFF);  // OFF1, NOFF1, OFF2, NOFF2 of this lane
#pragma unroll
            for (int u = 0; u < N; ++u) {
                l[u][0] = p_add(H[u][0], NOE1); l[u][1] = p_addmax(l[u][0], NE1, p_add(H[u][1], NOE1)); l[u][2] = p_addmax(l[u][1], NE1, p_add(H[u][2], NOE1));
                const unsigned lout = p_addmax(l[u][2], NE1, p_add(H[u][3], NOE1));
                k_[u][0] = p_add(H[u][0], NOE2); k_[u][1] = p_addmax(k_[u][0], NE2, p_add(H[u][1], NOE2)); k_[u][2] = p_addmax(k_[u][1], NE2, p_add(H[u][2], NOE2));
                const unsigned kout = p_addmax(k_[u][2], NE2, p_add(H[u][3], NOE2));
                g1[u] = p_add(lout, lc4.x); g2[u] = p_add(kout, lc4.z);
            }
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
                for (int u = 0; u < N; ++u) {
                    const unsigned u1 = (unsigned)poa_shfl_up((int)g1[u], d), u2 = (unsigned)poa_shfl_up((int)g2[u], d);
                    g1[u] = p_max(g1[u], u1); g2[u] = p_max(g2[u], u2);  // lanes < d get their own value back: a no-op
                }
            }
            unsigned x1[N], x2[N], t1[N], t2[N];
#pragma unroll
            for (int u = 0; u < N; ++u) {
                x1[u] = (unsigned)poa_shfl_up((int)g1[u], 1); x2[u] = (unsigned)poa_shfl_up((int)g2[u], 1);
                t1[u] = (unsigned)poa_shfl((int)g1[u], 31); t2[u] = (unsigned)poa_shfl((int)g2[u], 31);
                if (lane == 0) { x1[u] = NEGLP; x2[u] = NEGLP; }
            }
            // ---- the scalar carry runs through the N scan totals; then fix-up, H, new E (abpoa_align_simd.c:1060-1071), stores
#pragma unroll
            for (int u = 0; u < N; ++u) {
                const unsigned xx1 = p_max3(x1[u], p_lolo(NEGLP, t1[u]), carry1);  // high halves continue after all low halves
                const unsigned xx2 = p_max3(x2[u], p_lolo(NEGLP, t2[u]), carry2);
                const unsigned fin1 = p_add(xx1, lc4.y), fin2 = p_add(xx2, lc4.w);
                carry1 = p_add(p_max3(t1[u], p_swap(t1[u]), carry1), NCW1);
                carry2 = p_add(p_max3(t2[u], p_swap(t2[u]), carry2), NCW2);
                const unsigned F10 = fin1, F11 = p_addmax(fin1, NE1, l[u][0]), F12 = p_addmax(fin1, NE1_2, l[u][1]), F13 = p_addmax(fin1, NE1_3, l[u][2]);
                const unsigned F20 = fin2, F21 = p_addmax(fin2, NE2, k_[u][0]), F22 = p_addmax(fin2, NE2_2, k_[u][1]), F23 = p_addmax(fin2, NE2_3, k_[u][2]);
                H[u][0] = p_max3(H[u][0], F10, F20); H[u][1] = p_max3(H[u][1], F11, F21); H[u][2] = p_max3(H[u][2], F12, F22); H[u][3] = p_max3(H[u][3], F13, F23);
                if (local) { H[u][0] = p_max(H[u][0], ZERO); H[u][1] = p_max(H[u][1], ZERO); H[u][2] = p_max(H[u][2], ZERO); H[u][3] = p_max(H[u][3], ZERO); }
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    A[u][x] = p_addmax(A[u][x], NE1, p_add(H[u][x], NOE1));
                    B[u][x] = p_addmax(B[u][x], NE2, p_add(H[u][x], NOE2));
                    if (local) { A[u][x] = p_max(A[u][x], ZERO); B[u][x] = p_max(B[u][x], ZERO); }
                }
            }
            if (bnd) {
#pragma unroll
                for (int u = 0; u < N; ++u) {
                    if (u != 0 && u != N - 1) continue;
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        H[u][x] = (H[u][x] & ~m[u][x]) | (INFP & m[u][x]);
                        A[u][x] = (A[u][x] & ~m[u][x]) | (INFP & m[u][x]);
                        B[u][x] = (B[u][x] & ~m[u][x]) | (INFP & m[u][x]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < N; ++u) {
                p16_st(dst, H[u][0], H[u][1], H[u][2], H[u][3]);
                p16_st(dst + pstride, A[u][0], A[u][1], A[u][2], A[u][3]);
                p16_st(dst + 2 * pstride, B[u][0], B[u][1], B[u][2], B[u][3]);
                dst += P16_CPB;
                if (cur_res) {  // after every predecessor read of this pass: a lane only ever touches its own slices
                    ring_st(ring, rslot, H[u][0], H[u][1], H[u][2], H[u][3]); ring_st(ring, rslot + P16_CPB, A[u][0], A[u][1], A[u][2], A[u][3]);
                    ring_st(ring, rslot + 2 * P16_CPB, B[u][0], B[u][1], B[u][2], B[u][3]);
                }
                rslot = rslot == (P16_SMCH - 1) * (3 * P16_CPB) ? 0u : rslot + 3 * P16_CPB;
            }
            // ---- row maximum (abpoa_align_simd.c:1107-1119): remember the first and the last chunk attaining it
            if (track) {
                int cmx[N];
#pragma unroll
                for (int u = 0; u < N; ++u) {
                    const unsigned cm = p_max(p_max3(H[u][0], H[u][1], H[u][2]), H[u][3]);
                    cmx[u] = poa_redux_max(imax(p_lo(cm), p_hi(cm)));
                }
#pragma unroll
                for (int u = 0; u < N; ++u) {
                    if (cmx[u] > rmx) { rmx = cmx[u]; fc = c + u; }
                    if (cmx[u] >= rmx) lc = c + u;
                }
            }
        };
        auto row_passes = [&](auto staged) {
            int c = cb;
#pragma unroll 1
            for (; c + (POA_P16_ILP - 1) <= ce; c += POA_P16_ILP) pass(p16_n<POA_P16_ILP>(), staged, c);
#if POA_P16_ILP >= 4
            if (c + 2 <= ce) { pass(p16_n<3>(), staged, c); c += 3; }
#endif
#if POA_P16_ILP >= 3
            if (c + 1 <= ce) { pass(p16_n<2>(), staged, c); c += 2; }
#endif
#if POA_P16_ILP >= 2
            if (c <= ce) pass(p16_n<1>(), staged, c);
#endif
        };
        if (qst) row_passes(p16_b<true>()); else row_passes(p16_b<false>());
        prev_meta = poa_make_int4((int)roff, beg, end, 0);
        prev_res = cur_res;
        if (lane == 0) rowmeta[i] = prev_meta;
        if (track) {
            // first / last column holding the row maximum: bit r of a lane's mask = low-half cell r equals it, bit 4+r = high half
            const unsigned pat = p_pack(rmx, rmx);
            // H of the first / last chunk holding the maximum: from the ring, else from the slab (one chunk for most rows)
            auto eqmask = [&](const int ch) {
                const uint4 h = cur_res ? ring_ld(ring, (unsigned)(ch % P16_SMCH) * (3 * P16_CPB)) : p16_ld(slab_lane + (size_t)(roff + (unsigned)(ch - cb)) * P16_CPB);
                const unsigned z = p_minu(h.x ^ pat, 0x00010001u) | (p_minu(h.y ^ pat, 0x00010001u) << 1)
                                 | (p_minu(h.z ^ pat, 0x00010001u) << 2) | (p_minu(h.w ^ pat, 0x00010001u) << 3);
                return (~z & 0xfu) | ((~z >> 12) & 0xf0u);
            };
            const unsigned fm = eqmask(fc);
            unsigned lm = fm;
            if (lc != fc) lm = eqmask(lc);
            int first = INT_MAX, last = -1;
            if (fm) { const int b = p_ctz(fm); first = fc * P16_CW + lane * 4 + b + (b >= 4 ? 124 : 0); }
            if (lm) { const int b = 31 - p_clz(lm); last = lc * P16_CW + lane * 4 + b + (b >= 4 ? 124 : 0); }
            const int left = poa_redux_min(first), right = poa_redux_max(last);
            prev_left = left; prev_right = right;
            if (local && rmx > best_score) { best_score = rmx; best_i = i; best_j = left; }  // abpoa_align_simd.c:1208-1210
            if (wb >= 0) {  // abpoa_align_simd.c:1121-1130; reductions without a return value: nothing to wait for
                cpa_wait_pending(0);
                if (lane < ri.w) { const int out_row = ring_ld32(sm, P16_OUTS_OFF + lane * 4); poa_red_max(&mprr[out_row], right + 1); poa_red_min(&mplr[out_row], left + 1); }
#pragma unroll 1
                for (int k = lane + POA_WARP; k < ri.w; k += POA_WARP) {
                    const int o = pool_row[ri.z + k];
                    poa_red_max(&mprr[o], right + 1); poa_red_min(&mplr[o], left + 1);
                }
            }
        }
        // no barrier here: the next row starts with one (after its cp.async wait)
    }
    cpa_wait_pending(0);
    sync_block<NW>();
    // ---- global best (abpoa_align_simd.c:1092-1105)
    if (lane == 0) {
        if (!local) {
            const int4 ri = rowinfo[rows];  // sink row
            for (int k = 0; k < ri.y; ++k) {
                const int pi = pool_row[ri.x + k];
                const int4 pm = rowmeta[pi];
                const int e = qlen > pm.z ? pm.z : qlen;
                const int sc = *cell_ptr16(sh.ws, pm, 0, e);
                if (sc > best_score) { best_score = sc; best_i = pi; best_j = e; }
            }
        }
        sh.best_score = best_score; sh.best_i = best_i; sh.best_j = best_j;
        sh.inband += inband; sh.edge_rows += edge_rows;
    }
    sync_block<NW>();
}
#undef INFP
#undef NEGLP
#undef NOE1
#undef NOE2
#undef NE1
#undef NE2
#undef NE1_2
#undef NE1_3
#undef NE2_2
#undef NE2_3
#undef NCW1
#undef NCW2
#undef P16_PK

// ------------------------------------------------------------------------------------------------
// F1 / F2 of row i at columns j and j - 1, for the traceback's insertion steps (abpoa_align_simd.c:420-445 reads
// dp_f1[j], dp_f1[j-1], dp_f2[j], dp_f2[j-1]).  The fill does not store the F planes; this re-runs its arithmetic --
// the same packed operations in the same order -- for chunks beg >> 8 .. j >> 8 of that one row: predecessors' H / E1 / E2
// from the slab, the profile from qp[] (both still in place while the alignment is traced back).  One warp; every lane
// returns the same four values {F1[j], F2[j], F1[j-1], F2[j-1]}; the j - 1 entries are inf_min when j - 1 < beg, which is
// what the traceback substitutes there.  Cost: one row (3-4 chunks) per insertion step, about 0.1 % of the fill's work.
POA_DN void p16_row_f(Shared &sh, const DevParams &P, int qlen, int i, int j, int *out) {
    Ws &w = sh.ws;
    const int lane = poa_tid() % POA_WARP;
    const int inf_min = inf_min_of<short>(P);
    const bool local = P.local != 0;
    const int e1 = P.e1, e2 = P.e2, oe1 = P.oe1, oe2 = P.oe2;
    const int emax = imax(e1, e2);
    const int negl = -32768 + 8 * emax + 8;
    const unsigned INFP = p_pack(inf_min, inf_min), NEGLP = p_pack(negl, negl);
    const unsigned NOE1 = p_pack(-oe1, -oe1), NOE2 = p_pack(-oe2, -oe2), NE1 = p_pack(-e1, -e1), NE2 = p_pack(-e2, -e2);
    const unsigned NE1_2 = p_add(NE1, NE1), NE1_3 = p_add(NE1_2, NE1);
    const unsigned NE2_2 = p_add(NE2, NE2), NE2_3 = p_add(NE2_2, NE2);
    const unsigned OFF1 = p_pack(e1 * 4 * (lane + 1), e1 * 4 * (lane + 33)), NOFF1 = p_pack(-e1 * 4 * lane, -e1 * 4 * (lane + 32));
    const unsigned OFF2 = p_pack(e2 * 4 * (lane + 1), e2 * 4 * (lane + 33)), NOFF2 = p_pack(-e2 * 4 * lane, -e2 * 4 * (lane + 32));
    const unsigned NCW1 = p_pack(-e1 * P16_CW, -e1 * P16_CW), NCW2 = p_pack(-e2 * P16_CW, -e2 * P16_CW);
    const int f0_1 = imax(inf_min - oe1, inf_min - e1), f0_2 = imax(inf_min - oe2, inf_min - e2);
    const int nchq = (qlen >> 8) + 1;
    const int4 rm = w.rowmeta[i], ri = w.rowinfo[i];
    const int beg = rm.y, end = rm.z, cb = beg >> 8, ce = end >> 8, rb = w.rbase[i];
    const char *slab = w.slab, *slab_lane = slab + lane * 16;
    const char *qrow = w.qp + (size_t)(unsigned)(rb * nchq) * P16_CPB + lane * 16;
    unsigned carry1 = p_pack(f0_1 + e1 * (beg - cb * P16_CW), f0_1 + e1 * (beg - cb * P16_CW));
    unsigned carry2 = p_pack(f0_2 + e2 * (beg - cb * P16_CW), f0_2 + e2 * (beg - cb * P16_CW));
    int res[4] = {inf_min, inf_min, inf_min, inf_min};
    const int cj = j >> 8;
    for (int c = cb; c <= cj; ++c) {
        const int c0 = c * P16_CW;
        unsigned M0 = INFP, M1 = INFP, M2 = INFP, M3 = INFP, A0 = INFP, A1 = INFP, A2 = INFP, A3 = INFP, B0 = INFP, B1 = INFP, B2 = INFP, B3 = INFP;
        for (int k = 0; k < ri.y; ++k) {  // predecessors (abpoa_align_simd.c:966-1029)
            const int pk = w.pool_row[ri.x + k];
            const int4 pm = w.rowmeta[pk];
            const int pcb = pm.y >> 8, pce = pm.z >> 8;
            if (c >= pcb && c <= pce + 1) {
                const unsigned pn_ = (unsigned)(pce - pcb + 1), idx = (unsigned)pm.x + (unsigned)(c - pcb);
                int prevlast = inf_min;
                if (c > pcb) prevlast = *reinterpret_cast<const short *>(slab + (size_t)idx * P16_CPB - 2);
                if (c <= pce) {
                    const uint4 h = p16_ld(slab_lane + (size_t)idx * P16_CPB);
                    const uint4 a = p16_ld(slab_lane + (size_t)(idx + pn_) * P16_CPB);
                    const uint4 b = p16_ld(slab_lane + (size_t)(idx + 2 * pn_) * P16_CPB);
                    const unsigned rot = (unsigned)poa_shfl((int)h.w, (lane + 31) & 31);
                    const unsigned s0 = lane == 0 ? p_pack(prevlast, p_lo(rot)) : rot;
                    M0 = p_max(M0, s0); M1 = p_max(M1, h.x); M2 = p_max(M2, h.y); M3 = p_max(M3, h.z);
                    A0 = p_max(A0, a.x); A1 = p_max(A1, a.y); A2 = p_max(A2, a.z); A3 = p_max(A3, a.w);
                    B0 = p_max(B0, b.x); B1 = p_max(B1, b.y); B2 = p_max(B2, b.z); B3 = p_max(B3, b.w);
                } else if (lane == 0) {
                    M0 = p_max(M0, p_pack(prevlast, inf_min));
                }
            }
        }
        if (local && c == 0 && lane == 0) M0 = p_max(M0, p_pack(0, inf_min));
        const uint4 qv = p16_ld(qrow + (size_t)(unsigned)c * P16_CPB);
        unsigned H0 = p_max3(p_add(M0, qv.x), A0, B0), H1 = p_max3(p_add(M1, qv.y), A1, B1);
        unsigned H2 = p_max3(p_add(M2, qv.z), A2, B2), H3 = p_max3(p_add(M3, qv.w), A3, B3);
        if ((c == cb && beg > c0) || (c == ce && end < c0 + P16_CW - 1)) {
            const int brel = imax(beg - c0, 0), erel = imin(end - c0, P16_CW - 1);
            const unsigned da = p_pack(lane * 4 - brel, 128 + lane * 4 - brel), db = p_pack(erel - lane * 4, erel - 128 - lane * 4);
            const unsigned m0 = p_signmask(p_min(da, db));
            const unsigned m1 = p_signmask(p_min(p_add(da, 0x00010001u), p_add(db, 0xffffffffu)));
            const unsigned m2 = p_signmask(p_min(p_add(da, 0x00020002u), p_add(db, 0xfffefffeu)));
            const unsigned m3 = p_signmask(p_min(p_add(da, 0x00030003u), p_add(db, 0xfffdfffdu)));
            H0 = (H0 & ~m0) | (INFP & m0); H1 = (H1 & ~m1) | (INFP & m1);
            H2 = (H2 & ~m2) | (INFP & m2); H3 = (H3 & ~m3) | (INFP & m3);
        }
        unsigned F1[4], F2[4];
        {
            const unsigned l1 = p_add(H0, NOE1), l2 = p_addmax(l1, NE1, p_add(H1, NOE1)), l3 = p_addmax(l2, NE1, p_add(H2, NOE1));
            const unsigned lout = p_addmax(l3, NE1, p_add(H3, NOE1));
            const unsigned k1 = p_add(H0, NOE2), k2 = p_addmax(k1, NE2, p_add(H1, NOE2)), k3 = p_addmax(k2, NE2, p_add(H2, NOE2));
            const unsigned kout = p_addmax(k3, NE2, p_add(H3, NOE2));
            unsigned g1 = p_add(lout, OFF1), g2 = p_add(kout, OFF2);
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned u1 = (unsigned)poa_shfl_up((int)g1, d), u2 = (unsigned)poa_shfl_up((int)g2, d);
                g1 = p_max(g1, u1); g2 = p_max(g2, u2);
            }
            unsigned x1 = (unsigned)poa_shfl_up((int)g1, 1), x2 = (unsigned)poa_shfl_up((int)g2, 1);
            const unsigned t1 = (unsigned)poa_shfl((int)g1, 31), t2 = (unsigned)poa_shfl((int)g2, 31);
            if (lane == 0) { x1 = NEGLP; x2 = NEGLP; }
            x1 = p_max3(x1, p_lolo(NEGLP, t1), carry1);
            x2 = p_max3(x2, p_lolo(NEGLP, t2), carry2);
            const unsigned fin1 = p_add(x1, NOFF1), fin2 = p_add(x2, NOFF2);
            carry1 = p_add(p_max3(t1, p_swap(t1), carry1), NCW1);
            carry2 = p_add(p_max3(t2, p_swap(t2), carry2), NCW2);
            F1[0] = fin1; F1[1] = p_addmax(fin1, NE1, l1); F1[2] = p_addmax(fin1, NE1_2, l2); F1[3] = p_addmax(fin1, NE1_3, l3);
            F2[0] = fin2; F2[1] = p_addmax(fin2, NE2, k1); F2[2] = p_addmax(fin2, NE2_2, k2); F2[3] = p_addmax(fin2, NE2_3, k3);
        }
        for (int t = 0; t < 2; ++t) {  // columns j and j - 1 that fall into this chunk
            const int x = j - t;
            if ((x >> 8) != c || x < beg) continue;
            const int u = x & 255, src = (u & 127) >> 2, r = u & 3;
            const unsigned f1 = r == 0 ? F1[0] : r == 1 ? F1[1] : r == 2 ? F1[2] : F1[3];
            const unsigned f2 = r == 0 ? F2[0] : r == 1 ? F2[1] : r == 2 ? F2[2] : F2[3];
            const int v1 = (u >> 7) ? p_hi(f1) : p_lo(f1), v2 = (u >> 7) ? p_hi(f2) : p_lo(f2);
            res[2 * t] = poa_shfl(v1, src); res[2 * t + 1] = poa_shfl(v2, src);
        }
    }
    out[0] = res[0]; out[1] = res[1]; out[2] = res[2]; out[3] = res[3];
}

// ------------------------------------------------------------------------------------------------
// Packed 16-bit fill for NW > 1 warps per POA block (small or deep batches: few blocks, long rows -- with one block per SM the
// only thing that matters is the latency of one row, and the chunks of a row are the only parallelism a row has).
// Same arithmetic and row layout as fill_p16(); the 256-column chunks of a row are dealt to the warps (chunk c belongs to warp
// c % NW), NW chunks per round:
//   row start  -- barrier --  every warp reads the row's metadata from shared memory (gathered one row ahead by warp 0 with
//              cp.async, exactly as in fill_p16: rowinfo, the first two predecessors' descriptors, base, band inputs) and the
//              previous row's arg-max candidates, and computes the band redundantly (a few dozen ALU instructions);
//   phase A    each warp gathers its chunk's predecessors, runs the per-lane chains and the lane-half scan; lane 0 publishes
//              the chunk's scan total (per gap piece) in shared memory                                   -- barrier --
//   phase B    every warp replays the cheap scalar carry chain over the round's totals up to its own chunk, fixes up F,
//              finishes H / E, stores; after its last chunk it publishes its own row maximum with the first / last column
//              attaining it (from its ring slots: no global re-read), which the next row start combines.
// One barrier per row plus one per round, no global-memory round trip on the row-to-row critical path.  A warp only ever
// reads ring slots it wrote itself (chunk ownership is by absolute chunk number); the one cross-warp operand, the
// predecessor's last cell of the previous chunk, goes through a small shared array.
// ------------------------------------------------------------------------------------------------
constexpr int P16_MW_SMCH = 3;  // ring chunks per warp (rows up to 3 * NW chunks stay resident)
template <int NW> struct p16_mw_smem {
    static constexpr int o_xch = NW * P16_MW_SMCH * 3 * P16_CPB;  // [2 round parities][NW][2] scan totals
    static constexpr int o_rowx = o_xch + 2 * NW * 2 * 4;         // [2 row parities][NW][4] row maximum, first / last column
    static constexpr int o_lastH = o_rowx + 2 * NW * 4 * 4;       // [2 row parities][64] last H cell of a chunk
    static constexpr int o_meta = (o_lastH + 2 * 64 * 4 + 15) & ~15;  // 2 row metadata slots (P16_MW_SLOT), 2 x 128 B successor rows
    static constexpr int bytes = o_meta + 2 * P16_MW_SLOT + 2 * 128;
};
template <int NW> constexpr int p16_mw_smem_bytes() { return p16_mw_smem<NW>::bytes; }

template <int NW, bool LOCAL>
POA_DN void fill_p16_mw(Shared &sh, const DevParams &P, const uint8_t *q, int qlen, long long slab_bytes, const int pn) {
    Ws &w = sh.ws;
    const int tid = poa_tid(), lane = tid & 31, wid = tid >> 5;
    const int n_node = sh.n_node;
    const int rows = n_node - 1;
    const int inf_min = inf_min_of<short>(P);
    constexpr bool local = LOCAL;
    const int wb = local ? -1 : P.wb;
#ifdef POA_HOST_EMU
    const int bw = wb < 0 ? qlen : wb + (int)(P.wf * qlen);
#else
    const int bw = wb < 0 ? qlen : wb + (int)__fmul_rn(P.wf, (float)qlen);
#endif
    const int e1 = P.e1, e2 = P.e2, oe1 = P.oe1, oe2 = P.oe2;
    char *const slab = w.slab, *const qp = w.qp;
    const int4 *const rowinfo = w.rowinfo;
    int4 *const rowmeta = w.rowmeta;
    const int *const pool_row = w.pool_row, *const rr = w.rr;
    const int4 *const pred4 = w.pred4;
    int *const mplr = w.mplr, *const mprr = w.mprr;
    const uint8_t *const rbase = w.rbase;
#ifndef POA_HOST_EMU
    __builtin_assume(__isGlobal(slab)); __builtin_assume(__isGlobal(qp)); __builtin_assume(__isGlobal(rowinfo));
    __builtin_assume(__isGlobal(rowmeta)); __builtin_assume(__isGlobal(pool_row)); __builtin_assume(__isGlobal(rr));
    __builtin_assume(__isGlobal(mplr)); __builtin_assume(__isGlobal(mprr)); __builtin_assume(__isGlobal(rbase));
    __builtin_assume(__isGlobal(q)); __builtin_assume(__isGlobal(pred4));
#endif
    const long long slab_units = slab_bytes / P16_CPB;
    long long used = 0, inband = 0, edge_rows = 0;
    const int emax = imax(e1, e2);
    const int negl = -32768 + 8 * emax + 8;
    const unsigned INFP = p_pack(inf_min, inf_min), NEGLP = p_pack(negl, negl), ZERO = 0u;
    const unsigned NOE1 = p_pack(-oe1, -oe1), NOE2 = p_pack(-oe2, -oe2), NE1 = p_pack(-e1, -e1), NE2 = p_pack(-e2, -e2);
    const unsigned NE1_2 = p_add(NE1, NE1), NE1_3 = p_add(NE1_2, NE1), NE2_2 = p_add(NE2, NE2), NE2_3 = p_add(NE2_2, NE2);
    const unsigned OFF1 = p_pack(e1 * 4 * (lane + 1), e1 * 4 * (lane + 33)), NOFF1 = p_pack(-e1 * 4 * lane, -e1 * 4 * (lane + 32));
    const unsigned OFF2 = p_pack(e2 * 4 * (lane + 1), e2 * 4 * (lane + 33)), NOFF2 = p_pack(-e2 * 4 * lane, -e2 * 4 * (lane + 32));
    const int f0_1 = imax(inf_min - oe1, inf_min - e1), f0_2 = imax(inf_min - oe2, inf_min - e2);
    // shared memory: per-warp rings, round totals (two parities), per-warp row maxima (two row parities), last H cell per chunk
    // (two row parities), metadata slots
    typedef p16_mw_smem<NW> SM;
    const ring_ptr_t ring = ring_base(sh.ring + wid * (P16_MW_SMCH * 3 * P16_CPB), lane);
    int *const xch = reinterpret_cast<int *>(sh.ring + SM::o_xch);
    int *const rowx = reinterpret_cast<int *>(sh.ring + SM::o_rowx);
    int *const lastH = reinterpret_cast<int *>(sh.ring + SM::o_lastH);
    const ring_ptr_t sm = ring_base(sh.ring, 0);
    constexpr unsigned META = (unsigned)SM::o_meta, OUTS = META + 2 * P16_MW_SLOT;

    const int nchq = (qlen >> 8) + 1;
    for (int bc = wid; bc < 5 * nchq; bc += NW) {  // query profile, chunked layout
        const int b = bc / nchq, c = bc % nchq;
        unsigned v[4];
        for (int r = 0; r < 4; ++r) {
            const int jl = c * P16_CW + lane * 4 + r, jh = jl + 128;
            const int sl = (jl == 0 || jl > qlen) ? 0 : P.mat[b * 5 + q[jl - 1]];
            const int s2 = jh > qlen ? 0 : P.mat[b * 5 + q[jh - 1]];
            v[r] = p_pack(sl, s2);
        }
        p16_st(qp + (size_t)(unsigned)bc * P16_CPB + lane * 16, v[0], v[1], v[2], v[3]);
    }
    {   // row 0
        int end0;
        if (wb >= 0) {
            if (tid == 0) {
                mplr[0] = 0; mprr[0] = 0;
                const int4 ri = rowinfo[0];
                for (int k = 0; k < ri.w; ++k) { int o = pool_row[ri.z + k]; mplr[o] = 1; mprr[o] = 1; }
            }
            end0 = imin(qlen, imax(0, rr[0]) + bw);
        } else end0 = qlen;
        const int nch = (end0 >> 8) + 1;
        if ((long long)P16_PLANES * nch > slab_units) { if (tid == 0) sh.err = ST_ESLAB; sync_block<NW>(); return; }
        if (tid == 0) rowmeta[0] = poa_make_int4(0, 0, end0, 0);
        for (int c = wid; c < nch; c += NW) {
            unsigned v[5][4];
            for (int r = 0; r < 4; ++r) {
                int x[2][5];
                for (int hf = 0; hf < 2; ++hf) {
                    const int j = c * P16_CW + hf * 128 + lane * 4 + r;
                    int *y = x[hf];
                    if (j > end0) { y[0] = y[1] = y[2] = y[3] = y[4] = inf_min; }
                    else if (local) { y[0] = y[1] = y[2] = y[3] = y[4] = 0; }
                    else if (j == 0) { y[0] = 0; y[1] = -oe1; y[2] = -oe2; y[3] = y[4] = inf_min; }
                    else { y[3] = -P.o1 - e1 * j; y[4] = -P.o2 - e2 * j; y[0] = imax((int)(short)y[3], (int)(short)y[4]); y[1] = y[2] = inf_min; }
                }
                for (int p = 0; p < 5; ++p) v[p][r] = p_pack(x[0][p], x[1][p]);
            }
            for (int p = 0; p < 3; ++p)  // H, E1, E2 (row 0 is never a traceback row: the walk stops at i == 0)
                p16_st(slab + ((long long)p * nch + c) * P16_CPB + lane * 16, v[p][0], v[p][1], v[p][2], v[p][3]);
        }
        used = (long long)P16_PLANES * nch;
        inband += end0 + 1;
        sync_block<NW>();  // also orders row 0's band inputs (thread 0) before the first gather reads them
    }
    int best_score = inf_min, best_i = 0, best_j = 0;
    const int pshift = 31 - p_clz((unsigned)pn);
    char *const slab_lane = slab + lane * 16;
    const bool track = local || wb >= 0;
    int4 prev_meta = rowmeta[0];
    int prev_left = 0, prev_right = 0;  // arg-max columns of row i - 1 ...
    int pp_left = INT_MAX - 1, pp_right = -1;  // ... and of row i - 2 (see the band computation)
    bool prev_res = false;
    int prev_out_z = 0, prev_out_n = 0;  // successor list of the previous row (band propagation, warp 0)
    // metadata gather (warp 0), as in fill_p16: ten 16-byte cells per slot, each the aligned 16-byte window of its array that
    // holds the item, fetched by ONE zero-filling copy instruction for the ten lanes.  +0 rowinfo, +16 / +80 / +112 / +128
    // rowmeta of the first four predecessors, +32 bases (16 rows), +48 mplr, +64 mprr, +144 rr (4 rows each), +96 pred4 of the next row.
    const char *gsrc = nullptr; unsigned gdst = 0; int gshift = 4, gmask = ~0;
    if (wid == 0) {
        if (lane == 0) { gsrc = (const char *)rowinfo; gdst = 0; }
        else if (lane == 1) { gsrc = (const char *)rowmeta; gdst = 16; }
        else if (lane == 2) { gsrc = (const char *)rbase; gdst = 32; gshift = 0; gmask = ~15; }
        else if (lane == 3) { gsrc = (const char *)pred4; gdst = 96; }
        else if (lane == 4 && wb >= 0) { gsrc = (const char *)rr; gdst = 144; gshift = 2; gmask = ~3; }
        else if (lane == 5 && wb >= 0) { gsrc = (const char *)mplr; gdst = 48; gshift = 2; gmask = ~3; }
        else if (lane == 6 && wb >= 0) { gsrc = (const char *)mprr; gdst = 64; gshift = 2; gmask = ~3; }
        else if (lane == 7) { gsrc = (const char *)rowmeta; gdst = 80; }
        else if (lane == 8) { gsrc = (const char *)rowmeta; gdst = 112; }
        else if (lane == 9) { gsrc = (const char *)rowmeta; gdst = 128; }
    }
    const int gsel = lane == 1 ? 1 : lane == 7 ? 2 : lane == 8 ? 3 : lane == 9 ? 4 : lane == 3 ? 5 : 0;  // whose row: n1, a predecessor, n1 + 1
    auto gather = [&](const int n1, const int4 &p4_n1, const int cur) {  // warp 0: row n1 into its slot
        if (n1 >= rows || gsrc == nullptr) return;
        const int idx = gsel == 1 ? p4_n1.x : gsel == 2 ? p4_n1.y : gsel == 3 ? p4_n1.z : gsel == 4 ? p4_n1.w : n1 + (gsel == 5 ? 1 : 0);
        // a predecessor's descriptor is fetched only if that row exists and is complete (else it is forwarded in registers)
        const bool live = (unsigned)idx < (unsigned)((gsel >= 1 && gsel <= 4) ? cur : rows);
        cpa16z(sm, META + (unsigned)(n1 & 1) * P16_MW_SLOT + gdst, gsrc + (live ? (unsigned)(idx & gmask) << gshift : 0u), live ? 16u : 0u);
    };
    int4 np4 = rows > 1 ? pred4[1] : poa_make_int4(0, -1, -1, -1);
    if (wid == 0) { gather(1, np4, 1); cpa_commit(); }
    int par = 0;
    // previous row's arg-max columns from the candidates every warp published (rowx, parity of that row)
    auto combine = [&](const int row) {
        // lane k < NW takes warp k's candidate {maximum, first, last column} (one 16-byte read; the other lanes repeat entry 0),
        // three warp reductions combine them: ~12 instructions on every warp's row-to-row chain instead of two loops over NW
        const uint4 e = ring_ld(sm, (unsigned)SM::o_rowx + (unsigned)(((row & 1) * NW + (lane < NW ? lane : 0)) * 16));
        const int gmx = poa_redux_max((int)e.x);
        const int left = poa_redux_min((int)e.x == gmx ? (int)e.y : INT_MAX), right = poa_redux_max((int)e.x == gmx ? (int)e.z : -1);
        prev_left = left; prev_right = right;
        if (local && gmx > best_score) { best_score = gmx; best_i = row; best_j = left; }  // abpoa_align_simd.c:1208-1210
    };

    for (int i = 1; i < rows; ++i) {
        if (wid == 0) cpa_wait_pending(0);
        sync_block<NW>();  // row start: metadata slot filled, previous row finished by every warp
        if (track && i > 1) {
            pp_left = prev_left; pp_right = prev_right;
            combine(i - 1);
            if (wb >= 0 && wid == 0) {  // abpoa_align_simd.c:1121-1130 for the previous row, before the next gather reads the band inputs
                if (lane < prev_out_n) { const int o = ring_ld32(sm, OUTS + (unsigned)((i - 1) & 1) * 128 + lane * 4); poa_red_max(&mprr[o], prev_right + 1); poa_red_min(&mplr[o], prev_left + 1); }
#pragma unroll 1
                for (int k = lane + POA_WARP; k < prev_out_n; k += POA_WARP) { const int o = pool_row[prev_out_z + k]; poa_red_max(&mprr[o], prev_right + 1); poa_red_min(&mplr[o], prev_left + 1); }
            }
        }
        const unsigned slot = META + (unsigned)(i & 1) * P16_MW_SLOT;
        const uint4 ri_u = ring_ld(sm, slot), m0_u = ring_ld(sm, slot + 16), m1_u = ring_ld(sm, slot + 80), m2_u = ring_ld(sm, slot + 112), m3_u = ring_ld(sm, slot + 128);
        const uint4 nn_u = ring_ld(sm, slot + 96);
        const int4 ri = poa_make_int4((int)ri_u.x, (int)ri_u.y, (int)ri_u.z, (int)ri_u.w);
        const int rb = (ring_ld32(sm, slot + 32 + 4u * (unsigned)((i >> 2) & 3)) >> (8 * (i & 3))) & 0xff;
        const int p0 = np4.x, s0 = np4.y, s2 = np4.z, s3 = np4.w;
        const int4 nnp4 = i + 1 < rows ? poa_make_int4((int)nn_u.x, (int)nn_u.y, (int)nn_u.z, (int)nn_u.w) : poa_make_int4(0, -1, -1, -1);
        const int r = ring_ld32(sm, slot + 144 + 4 * (i & 3));
        int ml = ring_ld32(sm, slot + 48 + 4 * (i & 3)), mr = ring_ld32(sm, slot + 64 + 4 * (i & 3));
        const int4 pm0 = p0 == i - 1 ? prev_meta : poa_make_int4((int)m0_u.x, (int)m0_u.y, (int)m0_u.z, (int)m0_u.w);
        const int4 pm1 = s0 == i - 1 ? prev_meta : poa_make_int4((int)m1_u.x, (int)m1_u.y, (int)m1_u.z, (int)m1_u.w);
        const int4 pm2 = s2 == i - 1 ? prev_meta : poa_make_int4((int)m2_u.x, (int)m2_u.y, (int)m2_u.z, (int)m2_u.w);
        const int4 pm3 = s3 == i - 1 ? prev_meta : poa_make_int4((int)m3_u.x, (int)m3_u.y, (int)m3_u.z, (int)m3_u.w);
        if (wid == 0) {
            gather(i + 1, nnp4, i);
            if (lane < ri.w) cpa4(sm, OUTS + (unsigned)(i & 1) * 128 + lane * 4, &pool_row[ri.z + lane]);
            cpa_commit();
        }
        np4 = nnp4;
        prev_out_z = ri.z; prev_out_n = ri.w;
        int beg, end;
        if (wb < 0) { beg = 0; end = qlen; }
        else {
            // ml / mr were fetched one row ago, when row i - 1 had not been evaluated and the reductions carrying row i - 2's
            // columns had only just been issued (by other lanes: nothing orders them before that fetch).  Both rows' contributions
            // are therefore forwarded in registers; min / max are idempotent, so a contribution that also arrived through memory
            // does no harm.  Rows further back had a whole row's time for their reductions to land.
            int min_pre_beg = pm0.y;
            bool from_prev = p0 == i - 1, from_pp = p0 == i - 2;
            if (ri.y > 1) { from_prev |= s0 == i - 1; from_pp |= s0 == i - 2; min_pre_beg = imin(min_pre_beg, pm1.y); }
            if (ri.y > 2) { from_prev |= s2 == i - 1; from_pp |= s2 == i - 2; min_pre_beg = imin(min_pre_beg, pm2.y); }
            if (ri.y > 3) { from_prev |= s3 == i - 1; from_pp |= s3 == i - 2; min_pre_beg = imin(min_pre_beg, pm3.y); }
#pragma unroll 1
            for (int k = 4; k < ri.y; ++k) {
                const int pk = pool_row[ri.x + k];
                from_prev |= pk == i - 1; from_pp |= pk == i - 2;
                min_pre_beg = imin(min_pre_beg, rowmeta[pk].y);
            }
            if (from_prev) { ml = imin(ml, prev_left + 1); mr = imax(mr, prev_right + 1); }
            if (from_pp && i > 2) { ml = imin(ml, pp_left + 1); mr = imax(mr, pp_right + 1); }
            beg = imax(0, imin(ml, r) - bw);
            end = imin(qlen, imax(mr, r) + bw);
            if ((beg >> pshift) < (min_pre_beg >> pshift)) beg = min_pre_beg;
        }
        if (end < beg) end = beg;
        const int cb = beg >> 8, ce = end >> 8, nch = ce - cb + 1;
        if (used + (long long)P16_PLANES * nch > slab_units) { if (tid == 0) sh.err = ST_ESLAB; sync_block<NW>(); return; }
        const unsigned roff = (unsigned)used;
        used += (long long)P16_PLANES * nch;
        inband += end - beg + 1;
        edge_rows += (long long)ri.y * (end - beg + 1);
        int cf1 = f0_1 + e1 * (beg - cb * P16_CW), cf2 = f0_2 + e2 * (beg - cb * P16_CW);  // F entering column cb*256
        int rmx = INT_MIN, fc = -1, lc = -1;  // this warp's chunks only
        const char *qrow = qp + (size_t)(unsigned)(rb * nchq) * P16_CPB + lane * 16;
        const size_t pstride = (size_t)(unsigned)nch * P16_CPB;
        const bool cur_res = nch <= P16_MW_SMCH * NW && nch <= 64;

#pragma unroll 1
        for (int cbase = cb; cbase <= ce; cbase += NW) {
            const int c = cbase + ((wid - (cbase % NW) + NW) % NW);  // the chunk of this round with c % NW == wid
            const bool act = c <= ce;
            const int c0 = c * P16_CW;
            const unsigned cslot = (unsigned)((c / NW) % P16_MW_SMCH) * (3 * P16_CPB);
            unsigned H0 = INFP, H1 = INFP, H2 = INFP, H3 = INFP, A0 = INFP, A1 = INFP, A2 = INFP, A3 = INFP, B0 = INFP, B1 = INFP, B2 = INFP, B3 = INFP;
            unsigned l1 = 0, l2 = 0, l3 = 0, k1 = 0, k2 = 0, k3 = 0, x1 = 0, x2 = 0, t1 = NEGLP, t2 = NEGLP, m0 = 0, m1 = 0, m2 = 0, m3 = 0;
            bool bnd = false;
            if (act) {
                unsigned M0 = INFP, M1 = INFP, M2 = INFP, M3 = INFP;
                // With one block per SM nothing hides a load but the block's own instruction stream, so everything this chunk
                // needs from memory is requested before any of it is consumed: the profile chunk and the chunks of the first TWO
                // predecessors (further ones, rare, take the loop below one at a time).
                const uint4 qv = p16_ld(qrow + (size_t)(unsigned)c * P16_CPB);
                auto fetch = [&](const int pk, const int4 &pm, uint4 &h, uint4 &a, uint4 &bb, int &prevlast, bool &inrange, bool &touch) {
                    const int pcb = pm.y >> 8, pce = pm.z >> 8;
                    touch = !(c < pcb || c > pce + 1);
                    inrange = touch && c <= pce;
                    prevlast = inf_min;  // H_p[c0 - 1]
                    if (!touch) return;
                    const unsigned pn_ = (unsigned)(pce - pcb + 1), idx = (unsigned)pm.x + (unsigned)(c - pcb);
                    const bool in_ring = prev_res && pk == i - 1;
                    if (c > pcb) prevlast = in_ring ? lastH[((i - 1) & 1) * 64 + ((c - 1) & 63)] : (int)*reinterpret_cast<const short *>(slab + (size_t)idx * P16_CPB - 2);
                    if (inrange) {
                        if (in_ring) { h = ring_ld(ring, cslot); a = ring_ld(ring, cslot + P16_CPB); bb = ring_ld(ring, cslot + 2 * P16_CPB); }
                        else {
                            h = p16_ld(slab_lane + (size_t)idx * P16_CPB);
                            a = p16_ld(slab_lane + (size_t)(idx + pn_) * P16_CPB);
                            bb = p16_ld(slab_lane + (size_t)(idx + 2 * pn_) * P16_CPB);
                        }
                    }
                };
                auto consume = [&](const uint4 &h, const uint4 &a, const uint4 &bb, const int prevlast, const bool inrange, const bool touch) {
                    if (!touch) return;
                    if (inrange) {
                        const unsigned rot = (unsigned)poa_shfl((int)h.w, (lane + 31) & 31);
                        const unsigned s0_ = lane == 0 ? p_pack(prevlast, p_lo(rot)) : rot;
                        M0 = p_max(M0, s0_); M1 = p_max(M1, h.x); M2 = p_max(M2, h.y); M3 = p_max(M3, h.z);
                        A0 = p_max(A0, a.x); A1 = p_max(A1, a.y); A2 = p_max(A2, a.z); A3 = p_max(A3, a.w);
                        B0 = p_max(B0, bb.x); B1 = p_max(B1, bb.y); B2 = p_max(B2, bb.z); B3 = p_max(B3, bb.w);
                    } else if (lane == 0) {
                        M0 = p_max(M0, p_pack(prevlast, inf_min));
                    }
                };
                {
                    uint4 h0 = {0, 0, 0, 0}, a0 = h0, b0 = h0, h1 = h0, a1 = h0, b1 = h0;
                    int pl0 = inf_min, pl1 = inf_min;
                    bool in0 = false, in1 = false, t0 = false, t1_ = false;
                    fetch(p0, pm0, h0, a0, b0, pl0, in0, t0);
                    if (ri.y > 1) fetch(s0, pm1, h1, a1, b1, pl1, in1, t1_);
                    // in in_id order (abpoa_align_simd.c:966-1029); the first predecessor's chunk is taken as it is (no max against -inf)
                    if (in0) {
                        const unsigned rot = (unsigned)poa_shfl((int)h0.w, (lane + 31) & 31);
                        M0 = lane == 0 ? p_pack(pl0, p_lo(rot)) : rot; M1 = h0.x; M2 = h0.y; M3 = h0.z;
                        A0 = a0.x; A1 = a0.y; A2 = a0.z; A3 = a0.w; B0 = b0.x; B1 = b0.y; B2 = b0.z; B3 = b0.w;
                    } else {
                        consume(h0, a0, b0, pl0, in0, t0);
                    }
                    if (ri.y > 1) consume(h1, a1, b1, pl1, in1, t1_);
                }
#pragma unroll 1
                for (int k = 2; k < ri.y; ++k) {
                    int pk = s2;
                    int4 pm = pm2;
                    if (k == 3) { pk = s3; pm = pm3; } else if (k > 3) { pk = pool_row[ri.x + k]; pm = rowmeta[pk]; }
                    uint4 h = {0, 0, 0, 0}, a = h, bb = h;
                    int pl = inf_min;
                    bool in_ = false, t_ = false;
                    fetch(pk, pm, h, a, bb, pl, in_, t_);
                    consume(h, a, bb, pl, in_, t_);
                }
                if (local && c == 0 && lane == 0) M0 = p_max(M0, p_pack(0, inf_min));
                H0 = p_max3(p_add(M0, qv.x), A0, B0); H1 = p_max3(p_add(M1, qv.y), A1, B1);
                H2 = p_max3(p_add(M2, qv.z), A2, B2); H3 = p_max3(p_add(M3, qv.w), A3, B3);
                bnd = (c == cb && beg > c0) || (c == ce && end < c0 + P16_CW - 1);
                if (bnd) {
                    const int brel = imax(beg - c0, 0), erel = imin(end - c0, P16_CW - 1);
                    const unsigned da = p_pack(lane * 4 - brel, 128 + lane * 4 - brel), db = p_pack(erel - lane * 4, erel - 128 - lane * 4);
                    m0 = p_signmask(p_min(da, db));
                    m1 = p_signmask(p_min(p_add(da, 0x00010001u), p_add(db, 0xffffffffu)));
                    m2 = p_signmask(p_min(p_add(da, 0x00020002u), p_add(db, 0xfffefffeu)));
                    m3 = p_signmask(p_min(p_add(da, 0x00030003u), p_add(db, 0xfffdfffdu)));
                    H0 = (H0 & ~m0) | (INFP & m0); H1 = (H1 & ~m1) | (INFP & m1);
                    H2 = (H2 & ~m2) | (INFP & m2); H3 = (H3 & ~m3) | (INFP & m3);
                }
                l1 = p_add(H0, NOE1); l2 = p_addmax(l1, NE1, p_add(H1, NOE1)); l3 = p_addmax(l2, NE1, p_add(H2, NOE1));
                const unsigned lout = p_addmax(l3, NE1, p_add(H3, NOE1));
                k1 = p_add(H0, NOE2); k2 = p_addmax(k1, NE2, p_add(H1, NOE2)); k3 = p_addmax(k2, NE2, p_add(H2, NOE2));
                const unsigned kout = p_addmax(k3, NE2, p_add(H3, NOE2));
                unsigned g1 = p_add(lout, OFF1), g2 = p_add(kout, OFF2);
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned u1 = (unsigned)poa_shfl_up((int)g1, d), u2 = (unsigned)poa_shfl_up((int)g2, d);
                    g1 = p_max(g1, u1); g2 = p_max(g2, u2);
                }
                x1 = (unsigned)poa_shfl_up((int)g1, 1); x2 = (unsigned)poa_shfl_up((int)g2, 1);
                t1 = (unsigned)poa_shfl((int)g1, 31); t2 = (unsigned)poa_shfl((int)g2, 31);
                if (lane == 0) { x1 = NEGLP; x2 = NEGLP; }
                x1 = p_max(x1, p_lolo(NEGLP, t1)); x2 = p_max(x2, p_lolo(NEGLP, t2));
            }
            if (lane == 0) {
                int *slot_ = xch + (par * NW + (c - cbase)) * 2;
                slot_[0] = imax(p_lo(t1), p_hi(t1)); slot_[1] = imax(p_lo(t2), p_hi(t2));
            }
            sync_block<NW>();
            // carry chain over the round's chunks: F entering each chunk's first column (scalar, F domain)
            int mine1 = cf1, mine2 = cf2;
            for (int k = 0; k < NW && cbase + k <= ce; ++k) {
                if (cbase + k == c) { mine1 = cf1; mine2 = cf2; }
                const int *slot_ = xch + (par * NW + k) * 2;
                cf1 = imax(slot_[0], cf1) - e1 * P16_CW; cf2 = imax(slot_[1], cf2) - e2 * P16_CW;
            }
            par ^= 1;
            if (act) {
                x1 = p_max(x1, p_pack(mine1, mine1)); x2 = p_max(x2, p_pack(mine2, mine2));
                const unsigned fin1 = p_add(x1, NOFF1), fin2 = p_add(x2, NOFF2);
                const unsigned F10 = fin1, F11 = p_addmax(fin1, NE1, l1), F12 = p_addmax(fin1, NE1_2, l2), F13 = p_addmax(fin1, NE1_3, l3);
                const unsigned F20 = fin2, F21 = p_addmax(fin2, NE2, k1), F22 = p_addmax(fin2, NE2_2, k2), F23 = p_addmax(fin2, NE2_3, k3);
                H0 = p_max3(H0, F10, F20); H1 = p_max3(H1, F11, F21); H2 = p_max3(H2, F12, F22); H3 = p_max3(H3, F13, F23);
                if (local) { H0 = p_max(H0, ZERO); H1 = p_max(H1, ZERO); H2 = p_max(H2, ZERO); H3 = p_max(H3, ZERO); }
                A0 = p_addmax(A0, NE1, p_add(H0, NOE1)); A1 = p_addmax(A1, NE1, p_add(H1, NOE1));
                A2 = p_addmax(A2, NE1, p_add(H2, NOE1)); A3 = p_addmax(A3, NE1, p_add(H3, NOE1));
                B0 = p_addmax(B0, NE2, p_add(H0, NOE2)); B1 = p_addmax(B1, NE2, p_add(H1, NOE2));
                B2 = p_addmax(B2, NE2, p_add(H2, NOE2)); B3 = p_addmax(B3, NE2, p_add(H3, NOE2));
                if (local) {
                    A0 = p_max(A0, ZERO); A1 = p_max(A1, ZERO); A2 = p_max(A2, ZERO); A3 = p_max(A3, ZERO);
                    B0 = p_max(B0, ZERO); B1 = p_max(B1, ZERO); B2 = p_max(B2, ZERO); B3 = p_max(B3, ZERO);
                }
                if (bnd) {
                    H0 = (H0 & ~m0) | (INFP & m0); H1 = (H1 & ~m1) | (INFP & m1); H2 = (H2 & ~m2) | (INFP & m2); H3 = (H3 & ~m3) | (INFP & m3);
                    A0 = (A0 & ~m0) | (INFP & m0); A1 = (A1 & ~m1) | (INFP & m1); A2 = (A2 & ~m2) | (INFP & m2); A3 = (A3 & ~m3) | (INFP & m3);
                    B0 = (B0 & ~m0) | (INFP & m0); B1 = (B1 & ~m1) | (INFP & m1); B2 = (B2 & ~m2) | (INFP & m2); B3 = (B3 & ~m3) | (INFP & m3);
                }
                char *dst = slab_lane + (size_t)(roff + (unsigned)(c - cb)) * P16_CPB;
                p16_st(dst, H0, H1, H2, H3); p16_st(dst + pstride, A0, A1, A2, A3); p16_st(dst + 2 * pstride, B0, B1, B2, B3);
                if (cur_res) {
                    ring_st(ring, cslot, H0, H1, H2, H3); ring_st(ring, cslot + P16_CPB, A0, A1, A2, A3); ring_st(ring, cslot + 2 * P16_CPB, B0, B1, B2, B3);
                    if (lane == 31) lastH[(i & 1) * 64 + (c & 63)] = p_hi(H3);
                }
                if (track) {
                    const unsigned cm = p_max(p_max3(H0, H1, H2), H3);
                    const int cmx = poa_redux_max(imax(p_lo(cm), p_hi(cm)));
                    if (cmx > rmx) { rmx = cmx; fc = c; }
                    if (cmx >= rmx) lc = c;
                }
            }
        }
        prev_meta = poa_make_int4((int)roff, beg, end, 0);
        prev_res = cur_res;
        if (tid == 0) rowmeta[i] = prev_meta;
        if (track) {
            // this warp's candidates: first / last column of ITS chunks holding ITS maximum; a lane only reads back what it wrote
            // itself (its ring slices, or its own 16-byte slices of the slab), so no barrier is needed here
            int first = INT_MAX, last = -1;
            if (fc >= 0) {
                const unsigned pat = p_pack(rmx, rmx);
                auto eqmask = [&](const int ch) {  // which of the lane's eight cells of chunk ch hold the maximum
                    const uint4 h = cur_res ? ring_ld(ring, (unsigned)((ch / NW) % P16_MW_SMCH) * (3 * P16_CPB)) : p16_ld(slab_lane + (size_t)(roff + (unsigned)(ch - cb)) * P16_CPB);
                    const unsigned z = p_minu(h.x ^ pat, 0x00010001u) | (p_minu(h.y ^ pat, 0x00010001u) << 1)
                                     | (p_minu(h.z ^ pat, 0x00010001u) << 2) | (p_minu(h.w ^ pat, 0x00010001u) << 3);
                    return (~z & 0xfu) | ((~z >> 12) & 0xf0u);
                };
                const unsigned fm = eqmask(fc);
                unsigned lm = fm;
                if (lc != fc) lm = eqmask(lc);  // a warp usually owns one chunk of a row
                if (fm) { const int b = p_ctz(fm); first = fc * P16_CW + lane * 4 + b + (b >= 4 ? 124 : 0); }
                if (lm) { const int b = 31 - p_clz(lm); last = lc * P16_CW + lane * 4 + b + (b >= 4 ? 124 : 0); }
            }
            first = poa_redux_min(first); last = poa_redux_max(last);
            if (lane == 0) { int *rx = rowx + ((i & 1) * NW + wid) * 4; rx[0] = rmx; rx[1] = first; rx[2] = last; }
        }
        // no barrier here: the next row starts with one
    }
    if (wid == 0) cpa_wait_pending(0);
    sync_block<NW>();
    if (track && rows > 1) {
        combine(rows - 1);  // the last row's maximum (local mode's best cell); its successors' band inputs are not needed any more
    }
    if (tid == 0) {  // global best (abpoa_align_simd.c:1092-1105)
        if (!local) {
            const int4 ri = rowinfo[rows];
            for (int k = 0; k < ri.y; ++k) {
                const int pi = pool_row[ri.x + k];
                const int4 pm = rowmeta[pi];
                const int e = qlen > pm.z ? pm.z : qlen;
                const int sc = *cell_ptr16(w, pm, 0, e);
                if (sc > best_score) { best_score = sc; best_i = pi; best_j = e; }
            }
        }
        sh.best_score = best_score; sh.best_i = best_i; sh.best_j = best_j;
        sh.inband += inband; sh.edge_rows += edge_rows;
    }
    sync_block<NW>();
}

#endif  // POA_WARP == 32
