// Host-side check of the compact 2-D inverse transforms (dav1d-mirror_b200/csrc/itx2.cuh
// compiled as plain C++, each phase run lane by lane) against the reference's own
// itxfm_add table from oracle/_ref/libdav1d_ref.so (reference src/itx_tmpl.c), for every
// populated (size, type) slot, 8/10/12 bit, dense blocks and packed coefficient boxes.
// TEST INFRASTRUCTURE ONLY.  Built and run by tests/test_host.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include "../../dav1d-mirror_b200/csrc/itx2.cuh"

typedef void (*itx8_fn)(uint8_t *, ptrdiff_t, int16_t *, int);
typedef void (*itx16_fn)(uint16_t *, ptrdiff_t, int32_t *, int, int);
struct ItxTable { void *fn[19][17]; };

static uint64_t s = 88172645463325252ull;
static uint32_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); }

template <typename pixel>
static void run_itx2(pixel *dst, int dstride, typename d1::PxTraits<pixel>::coef *cf, int tx, int txtp, int eob,
                     int cw4, int ch4, int bdmax, bool via_res)
{
    using namespace d1;
    const Itx2Blk b = itx2_setup<pixel>(tx, txtp, eob, cw4, ch4, bdmax);
    int G = b.sw > b.sh ? b.sw : b.sh;
    if (rnd() & 1) G = 32;
    static int tile[64 * 65];
    for (int i = 0; i < 64 * 65; i++) tile[i] = (int)rnd();   // stale shared memory
    // via_res: the residual goes to an int16 plane first and is added to the prediction afterwards
    // (what the intra pre-pass + executor do); else read-modify-write of dst (inter residuals)
    static int16_t resp[64 * 64];
    for (int i = 0; i < 64 * 64; i++) resp[i] = (int16_t)rnd();
    int16_t *res = via_res ? resp : nullptr;
    int dc = 0;
    if (b.dc_only) dc = itx2_dc_value<pixel>(b, cf);
    else {
        for (int gl = 0; gl < G; gl++) itx2_phase_stage<pixel>(b, gl, G, cf, tile, false);
        for (int gl = 0; gl < G; gl++) itx2_phase_rows<64>(b, gl, tile);
        for (int gl = 0; gl < G; gl++) itx2_phase_cols<64>(b, gl, G, tile);
    }
    for (int gl = 0; gl < G; gl++) itx2_phase_out<pixel>(b, gl, G, tile, dc, dst, dstride, res, b.w, bdmax);
    if (via_res)
        for (int y = 0; y < b.h; y++)
            for (int x = 0; x < b.w; x++)
                dst[y * dstride + x] = (pixel)clip_px<pixel>(dst[y * dstride + x] + resp[y * b.w + x], bdmax);
}

template <typename pixel>
static int check_slot(void *fn, int tx, int txtp, int bdmax, int iters) {
    using namespace d1;
    typedef typename PxTraits<pixel>::coef coef;
    const TxDim t = tx_dim(tx);
    const int w = t.w, h = t.h, sw = w < 32 ? w : 32, sh = h < 32 ? h : 32;
    const int cmax = sizeof(pixel) == 1 ? 32767 : bdmax > 1023 ? 524287 : 131071;
    const int stride = 80;
    for (int it = 0; it < iters; it++) {
        alignas(64) coef dense[32 * 32], dense_ref[32 * 32], packed[32 * 32];
        pixel a[64 * stride], b[64 * stride];
        for (int i = 0; i < 64 * stride; i++) a[i] = b[i] = (pixel)(rnd() & bdmax);
        memset(dense, 0, sizeof(dense));
        // non-zero box
        int bw = 4 * (1 + rnd() % (sw / 4)), bh = 4 * (1 + rnd() % (sh / 4));
        if (it % 5 == 0) { bw = sw; bh = sh; }
        if (it % 5 == 1) { bw = sw < 8 ? sw : 8; bh = sh < 8 ? sh : 8; }
        const int cls = it % 7;
        const int amp = cls < 2 ? 255 : cls < 4 ? bdmax * 8 : cls < 6 ? cmax / 4 : cmax;
        int eob = 0;
        if (it % 11 == 3) {
            dense[0] = (coef)((int)(rnd() % (2u * amp + 1)) - amp);      // dc only
            bw = bh = 4;
        } else {
            for (int x = 0; x < bw; x++)
                for (int y = 0; y < bh; y++) {
                    if (cls == 1 && (rnd() & 1)) continue;
                    int v = (int)(rnd() % (2u * amp + 1)) - amp;
                    if (txtp == 16) v = (int)(rnd() % (8u * bdmax + 1)) - 4 * bdmax;
                    if (it % 13 == 7) v = (rnd() & 1) ? amp : -amp;
                    dense[y + x * sh] = (coef)v;
                }
            eob = 1 + rnd() % (sw * sh - 1);
        }
        memcpy(dense_ref, dense, sizeof(dense));
        if (sizeof(pixel) == 1) ((itx8_fn)fn)((uint8_t *)a, stride, (int16_t *)dense_ref, eob);
        else ((itx16_fn)fn)((uint16_t *)a, stride * 2, (int32_t *)dense_ref, eob, bdmax);
        const bool use_dense = it % 3 == 0;
        if (use_dense) {
            run_itx2<pixel>(b, stride, dense, tx, txtp, eob, 0, 0, bdmax, (it & 1) != 0);
        } else {
            int k = 0;
            for (int x = 0; x < bw; x++)
                for (int y = 0; y < bh; y++) packed[k++] = dense[y + x * sh];
            run_itx2<pixel>(b, stride, packed, tx, txtp, eob, bw / 4, bh / 4, bdmax, (it & 1) != 0);
        }
        for (int y = 0; y < 64; y++)
            for (int x = 0; x < stride; x++) {
                                if (a[y * stride + x] != b[y * stride + x]) {
                    printf("tx %d txtp %d bdmax %x it %d box %dx%d dense %d: mismatch at (%d,%d) ref %d got %d\n", tx, txtp,
                           bdmax, it, bw, bh, (int)use_dense, x, y, (int)a[y * stride + x], (int)b[y * stride + x]);
                    return 1;
                }
            }
    }
    return 0;
}

int main(int argc, char **argv) {
    void *h = dlopen(argc > 1 ? argv[1] : "oracle/_ref/libdav1d_ref.so", RTLD_NOW);
    if (!h) { printf("dlopen: %s\n", dlerror()); return 2; }
    const int iters = argc > 2 ? atoi(argv[2]) : 60;
    typedef void (*init_fn)(ItxTable *, int);
    init_fn i8 = (init_fn)dlsym(h, "dav1d_itx_dsp_init_8bpc"), i16 = (init_fn)dlsym(h, "dav1d_itx_dsp_init_16bpc");
    if (!i8 || !i16) { printf("missing dav1d_itx_dsp_init\n"); return 2; }
    int bad = 0, slots = 0;
    for (int bpc = 8; bpc <= 12; bpc += 2) {
        ItxTable tb;
        memset(&tb, 0, sizeof(tb));
        if (bpc == 8) i8(&tb, 8); else i16(&tb, bpc);
        const int bdmax = (1 << bpc) - 1;
        for (int tx = 0; tx < 19; tx++)
            for (int txtp = 0; txtp < 17; txtp++) {
                if (!tb.fn[tx][txtp]) continue;
                slots++;
                bad |= bpc == 8 ? check_slot<uint8_t>(tb.fn[tx][txtp], tx, txtp, bdmax, iters)
                                : check_slot<uint16_t>(tb.fn[tx][txtp], tx, txtp, bdmax, iters);
            }
    }
    printf("itx2: %d slots x %d cases: %s\n", slots, iters, bad ? "FAIL" : "ok");
    return bad;
}
