// Host-side unit check of the register-resident 1-D inverse transforms
// (dav1d-mirror_b200/csrc/itx_1d.cuh compiled as plain C++) against the
// reference's own 1-D functions exported by oracle/_ref/libdav1d_ref.so
// (reference src/itx_1d.c).  Test infrastructure only.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include "../../dav1d-mirror_b200/csrc/itx_1d.cuh"

typedef void (*ref1d)(int32_t *, ptrdiff_t, int, int);
static uint64_t s = 88172645463325252ull;
static uint32_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); }

template <int N> static int run(void *h, const char *name, int kind, int nin) {
    ref1d f = (ref1d)dlsym(h, name);
    if (!f) { printf("missing %s\n", name); return 1; }
    int bad = 0;
    for (int it = 0; it < 20000 && !bad; it++) {
        const int mode = it % 4;
        const int lim = mode == 0 ? 32767 : mode == 1 ? (1 << 17) - 1 : mode == 2 ? (1 << 19) - 1 : 255;
        const d1::Clamp cl = { mode == 0 || mode == 3 ? -32768 : mode == 1 ? ~(0x3ff << 7) + 0 : (int)((unsigned)~0xfff << 7),
                               0 };
        d1::Clamp c2 = cl; c2.hi = ~c2.lo;
        int32_t a[64], b[64];
        for (int i = 0; i < 64; i++) {
            int v = (int)(rnd() % (2u * lim + 1)) - lim;
            if (i >= nin) v = 0;
            if ((it & 7) == 5 && (rnd() & 3)) v = 0;   // sparse
            if ((it & 15) == 9) v = (rnd() & 1) ? lim : -lim;  // extremes
            a[i] = b[i] = v;
        }
        if (N < 64) { for (int i = N; i < 64; i++) a[i] = b[i] = 0; }
        f(a, 1, c2.lo, c2.hi);
        int c[64]; memcpy(c, b, sizeof(c));
        d1::itx1d_run<N>(c, kind, c2);
        for (int i = 0; i < N; i++) if (a[i] != c[i]) { bad = 1; printf("%s mismatch it=%d i=%d ref=%d got=%d\n", name, it, i, a[i], c[i]); break; }
    }
    printf("%-28s %s\n", name, bad ? "FAIL" : "ok");
    return bad;
}

int main(int argc, char **argv) {
    void *h = dlopen(argc > 1 ? argv[1] : "oracle/_ref/libdav1d_ref.so", RTLD_NOW);
    if (!h) { printf("dlopen: %s\n", dlerror()); return 2; }
    int bad = 0;
    bad |= run<4>(h, "dav1d_inv_dct4_1d_c", d1::K_DCT, 4);
    bad |= run<8>(h, "dav1d_inv_dct8_1d_c", d1::K_DCT, 8);
    bad |= run<16>(h, "dav1d_inv_dct16_1d_c", d1::K_DCT, 16);
    bad |= run<32>(h, "dav1d_inv_dct32_1d_c", d1::K_DCT, 32);
    bad |= run<64>(h, "dav1d_inv_dct64_1d_c", d1::K_DCT, 32);
    bad |= run<4>(h, "dav1d_inv_adst4_1d_c", d1::K_ADST, 4);
    bad |= run<8>(h, "dav1d_inv_adst8_1d_c", d1::K_ADST, 8);
    bad |= run<16>(h, "dav1d_inv_adst16_1d_c", d1::K_ADST, 16);
    bad |= run<4>(h, "dav1d_inv_flipadst4_1d_c", d1::K_FLIPADST, 4);
    bad |= run<8>(h, "dav1d_inv_flipadst8_1d_c", d1::K_FLIPADST, 8);
    bad |= run<16>(h, "dav1d_inv_flipadst16_1d_c", d1::K_FLIPADST, 16);
    bad |= run<4>(h, "dav1d_inv_identity4_1d_c", d1::K_IDENTITY, 4);
    bad |= run<8>(h, "dav1d_inv_identity8_1d_c", d1::K_IDENTITY, 8);
    bad |= run<16>(h, "dav1d_inv_identity16_1d_c", d1::K_IDENTITY, 16);
    bad |= run<32>(h, "dav1d_inv_identity32_1d_c", d1::K_IDENTITY, 32);
    return bad;
}
