// TEST ONLY (CPU): the host-side parameter block of the POA engine (smoothxg_b200/csrc/poa_host.hpp: build_params, the
// counterpart of reference src/smooth.cpp:256-297 + deps/abPOA/src/abpoa_align.c:12-25,151-176).  Checks the score matrix, the
// packed constants of the 16-bit fill and when the default-scoring instantiation (fill_p16<.., PRESET>) is selected.
#define POA_HOST_EMU
#include <cstdio>
#include <cstring>
#include <string>
#include "../../smoothxg_b200/csrc/poa_host.hpp"
namespace poa { int set_err(int code, const std::string &) { return code; } }
using namespace poa;

static int fails = 0;
#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

static DevParams make(int m, int x, int o1, int e1, int o2, int e2, int mode, unsigned flags = 0) {
    poa_b200_params_t p; memset(&p, 0, sizeof(p));
    p.match = m; p.mismatch = x; p.gap_open1 = o1; p.gap_ext1 = e1; p.gap_open2 = o2; p.gap_ext2 = e2;
    p.align_mode = mode; p.wb = 311; p.wf = 0.03f; p.out_cons = 1;
    poa_b200_engine_opts_t o; memset(&o, 0, sizeof(o)); o.flags = flags;
    DevParams d; memset(&d, 0, sizeof(d));
    build_params(p, o, d);
    return d;
}
static unsigned pk(int v) { const unsigned h = (unsigned)v & 0xffffu; return h | (h << 16); }

int main() {
    {   // smoothxg's defaults (src/main.cpp:322-327): the preset instantiation, with exactly the literals it carries
        const DevParams d = make(1, 4, 6, 2, 26, 1, 0);
        CHECK(d.p16_ok); CHECK(d.p16_default);
        CHECK(d.mat[0] == 1 && d.mat[1] == -4 && d.mat[4] == 0 && d.mat[24] == 0);  // N scores 0 (abpoa_align.c:19-22)
        CHECK(d.oe1 == 8 && d.oe2 == 27 && d.gap_mode == 0);
        CHECK(d.pk_inf == pk(-31717) && d.pk_negl == pk(-32744) && d.pk_noe1 == pk(-8) && d.pk_noe2 == pk(-27));
        CHECK(d.pk_ne1_3 == pk(-6) && d.pk_ne2_3 == pk(-3) && d.pk_ncw1 == pk(-512) && d.pk_ncw2 == pk(-256));
        CHECK(d.pn16 == 32 && d.pn32 == 16);
    }
    {   // local mode keeps the preset (the flag is about scoring only)
        const DevParams d = make(1, 4, 6, 2, 26, 1, 1);
        CHECK(d.p16_default && d.local);
    }
    {   // any other scoring (the adaptive presets of src/smooth.cpp:2028-2062, a user's -p): generic instantiation
        CHECK(!make(1, 19, 39, 3, 81, 1, 0).p16_default);
        CHECK(!make(1, 9, 16, 2, 41, 1, 0).p16_default);
        CHECK(!make(2, 4, 6, 2, 26, 1, 0).p16_default);
        CHECK(!make(1, 4, 6, 2, 26, 2, 0).p16_default);
        CHECK(make(1, 9, 16, 2, 41, 1, 0).p16_ok);  // ... still on the packed fill
    }
    {   // affine / linear gaps (abpoa_align.c:87-91) and the "generic fill only" option bit: no packed fill, hence no preset
        CHECK(make(1, 4, 6, 2, 0, 0, 0).gap_mode == 1); CHECK(!make(1, 4, 6, 2, 0, 0, 0).p16_ok && !make(1, 4, 6, 2, 0, 0, 0).p16_default);
        CHECK(make(1, 4, 0, 2, 0, 0, 0).gap_mode == 2);
        CHECK(!make(1, 4, 6, 2, 26, 1, 0, 1u).p16_ok && !make(1, 4, 6, 2, 26, 1, 0, 1u).p16_default);
    }
    {   // the lane-count option of the band-start rule (flags bits 4-5)
        CHECK(make(1, 4, 6, 2, 26, 1, 0, 1u << 4).pn16 == 16); CHECK(make(1, 4, 6, 2, 26, 1, 0, 2u << 4).pn16 == 8);
    }
    if (fails) { printf("%d check(s) failed\n", fails); return 1; }
    printf("params ok\n");
    return 0;
}
