// 1-D inverse transforms of AV1 (DCT 4..64, ADST/flipADST 4..16, identity
// 4..32, WHT 4) on register-resident lanes.
//
// Every function works in place on `c[0], c[S], c[2S] ...` where S is a
// compile-time stride, so after inlining all indices are constants and the
// whole vector lives in registers.  The arithmetic (12-bit fixed-point
// rotations, clamp after each add/sub stage) follows the reference's
// src/itx_1d.c exactly; the two-term rotations are written through m12()/m11()
// with the constants already reduced mod 4096 where the reference writes
// `(c - 4096)` (itx_1d.c:40-63), which is the same value mod 2^32.
//   dct4  itx_1d.c:65-96     dct8  :98-149    dct16 :151-244
//   dct32 :246-434           dct64 :436-781
//   adst4 :783-802  adst8 :804-851  adst16 :853-962  flip :964-979
//   identity :983-1017       wht4 :1023-1038
#pragma once
#include "common.cuh"

namespace d1 {

struct Clamp {
    int lo, hi;
    // lo <= hi always: two min/max instead of the compare + select chain of the ternary form
    HD int operator()(int v) const {
#ifdef __CUDA_ARCH__
        return min(max(v, lo), hi);
#else
        return v < lo ? lo : v > hi ? hi : v;
#endif
    }
};

HD int m12(int a, int ca, int b, int cb) { return (a * ca + b * cb + 2048) >> 12; }
HD int m11(int a, int ca, int b, int cb) { return (a * ca + b * cb + 1024) >> 11; }
HD int r12(int a, int ca) { return (a * ca + 2048) >> 12; }
HD int s8(int v) { return (v * 181 + 128) >> 8; }

// ---------------------------------------------------------------- DCT
template <int S, bool TX64> HD void idct4(int *c, const Clamp cl) {
    const int i0 = c[0], i1 = c[S];
    int t0, t1, t2, t3;
    if (TX64) {
        t0 = t1 = s8(i0);
        t2 = r12(i1, 1567);
        t3 = r12(i1, 3784);
    } else {
        const int i2 = c[2 * S], i3 = c[3 * S];
        t0 = s8(i0 + i2);
        t1 = s8(i0 - i2);
        t2 = m12(i1, 1567, i3, 312) - i3;
        t3 = m12(i1, -312, i3, 1567) + i1;
    }
    c[0] = cl(t0 + t3);
    c[S] = cl(t1 + t2);
    c[2 * S] = cl(t1 - t2);
    c[3 * S] = cl(t0 - t3);
}

template <int S, bool TX64> HD void idct8(int *c, const Clamp cl) {
    idct4<2 * S, TX64>(c, cl);
    const int i1 = c[S], i3 = c[3 * S];
    int a4, a5, a6, a7;
    if (TX64) {
        a4 = r12(i1, 799);
        a5 = r12(i3, -2276);
        a6 = r12(i3, 3406);
        a7 = r12(i1, 4017);
    } else {
        const int i5 = c[5 * S], i7 = c[7 * S];
        a4 = m12(i1, 799, i7, 79) - i7;
        a5 = m11(i5, 1703, i3, -1138);
        a6 = m11(i5, 1138, i3, 1703);
        a7 = m12(i1, -79, i7, 799) + i1;
    }
    const int t4 = cl(a4 + a5);
    a5 = cl(a4 - a5);
    const int t7 = cl(a7 + a6);
    a6 = cl(a7 - a6);
    const int t5 = s8(a6 - a5), t6 = s8(a6 + a5);
    const int e0 = c[0], e1 = c[2 * S], e2 = c[4 * S], e3 = c[6 * S];
    c[0] = cl(e0 + t7);
    c[S] = cl(e1 + t6);
    c[2 * S] = cl(e2 + t5);
    c[3 * S] = cl(e3 + t4);
    c[4 * S] = cl(e3 - t4);
    c[5 * S] = cl(e2 - t5);
    c[6 * S] = cl(e1 - t6);
    c[7 * S] = cl(e0 - t7);
}

template <int S, bool TX64> HD void idct16(int *c, const Clamp cl) {
    idct8<2 * S, TX64>(c, cl);
    const int i1 = c[S], i3 = c[3 * S], i5 = c[5 * S], i7 = c[7 * S];
    int a8, a9, a10, a11, a12, a13, a14, a15;
    if (TX64) {
        a8 = r12(i1, 401);
        a9 = r12(i7, -2598);
        a10 = r12(i5, 1931);
        a11 = r12(i3, -1189);
        a12 = r12(i3, 3920);
        a13 = r12(i5, 3612);
        a14 = r12(i7, 3166);
        a15 = r12(i1, 4076);
    } else {
        const int i9 = c[9 * S], i11 = c[11 * S], i13 = c[13 * S], i15 = c[15 * S];
        a8 = m12(i1, 401, i15, 20) - i15;
        a9 = m11(i9, 1583, i7, -1299);
        a10 = m12(i5, 1931, i11, 484) - i11;
        a11 = m12(i13, -176, i3, -1189) + i13;
        a12 = m12(i13, 1189, i3, -176) + i3;
        a13 = m12(i5, -484, i11, 1931) + i5;
        a14 = m11(i9, 1299, i7, 1583);
        a15 = m12(i1, -20, i15, 401) + i1;
    }
    int t8 = cl(a8 + a9), t9 = cl(a8 - a9);
    int t10 = cl(a11 - a10), t11 = cl(a11 + a10);
    int t12 = cl(a12 + a13), t13 = cl(a12 - a13);
    int t14 = cl(a15 - a14), t15 = cl(a15 + a14);

    a9 = m12(t14, 1567, t9, 312) - t9;
    a14 = m12(t14, -312, t9, 1567) + t14;
    a10 = m12(t13, 312, t10, -1567) - t13;
    a13 = m12(t13, 1567, t10, 312) - t10;

    a8 = cl(t8 + t11);
    t9 = cl(a9 + a10);
    t10 = cl(a9 - a10);
    a11 = cl(t8 - t11);
    a12 = cl(t15 - t12);
    t13 = cl(a14 - a13);
    t14 = cl(a14 + a13);
    a15 = cl(t15 + t12);

    a10 = s8(t13 - t10);
    a13 = s8(t13 + t10);
    t11 = s8(a12 - a11);
    t12 = s8(a12 + a11);

    const int o[8] = { a15, t14, a13, t12, t11, a10, t9, a8 };
    int e[8];
#pragma unroll
    for (int k = 0; k < 8; k++) e[k] = c[2 * k * S];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        c[k * S] = cl(e[k] + o[k]);
        c[(15 - k) * S] = cl(e[k] - o[k]);
    }
}

template <int S, bool TX64> HD void idct32(int *c, const Clamp cl) {
    idct16<2 * S, TX64>(c, cl);
    const int i1 = c[S], i3 = c[3 * S], i5 = c[5 * S], i7 = c[7 * S];
    const int i9 = c[9 * S], i11 = c[11 * S], i13 = c[13 * S], i15 = c[15 * S];
    int a16, a17, a18, a19, a20, a21, a22, a23, a24, a25, a26, a27, a28, a29, a30, a31;
    if (TX64) {
        a16 = r12(i1, 201);
        a17 = r12(i15, -2751);
        a18 = r12(i9, 1751);
        a19 = r12(i7, -1380);
        a20 = r12(i5, 995);
        a21 = r12(i11, -2106);
        a22 = r12(i13, 2440);
        a23 = r12(i3, -601);
        a24 = r12(i3, 4052);
        a25 = r12(i13, 3290);
        a26 = r12(i11, 3513);
        a27 = r12(i5, 3973);
        a28 = r12(i7, 3857);
        a29 = r12(i9, 3703);
        a30 = r12(i15, 3035);
        a31 = r12(i1, 4091);
    } else {
        const int i17 = c[17 * S], i19 = c[19 * S], i21 = c[21 * S], i23 = c[23 * S];
        const int i25 = c[25 * S], i27 = c[27 * S], i29 = c[29 * S], i31 = c[31 * S];
        a16 = m12(i1, 201, i31, 5) - i31;
        a17 = m12(i17, -1061, i15, -2751) + i17;
        a18 = m12(i9, 1751, i23, 393) - i23;
        a19 = m12(i25, -239, i7, -1380) + i25;
        a20 = m12(i5, 995, i27, 123) - i27;
        a21 = m12(i21, -583, i11, -2106) + i21;
        a22 = m11(i13, 1220, i19, -1645);
        a23 = m12(i29, -44, i3, -601) + i29;
        a24 = m12(i29, 601, i3, -44) + i3;
        a25 = m11(i13, 1645, i19, 1220);
        a26 = m12(i21, 2106, i11, -583) + i11;
        a27 = m12(i5, -123, i27, 995) + i5;
        a28 = m12(i25, 1380, i7, -239) + i7;
        a29 = m12(i9, -393, i23, 1751) + i9;
        a30 = m12(i17, 2751, i15, -1061) + i15;
        a31 = m12(i1, -5, i31, 201) + i1;
    }
    int t16 = cl(a16 + a17), t17 = cl(a16 - a17);
    int t18 = cl(a19 - a18), t19 = cl(a19 + a18);
    int t20 = cl(a20 + a21), t21 = cl(a20 - a21);
    int t22 = cl(a23 - a22), t23 = cl(a23 + a22);
    int t24 = cl(a24 + a25), t25 = cl(a24 - a25);
    int t26 = cl(a27 - a26), t27 = cl(a27 + a26);
    int t28 = cl(a28 + a29), t29 = cl(a28 - a29);
    int t30 = cl(a31 - a30), t31 = cl(a31 + a30);

    a17 = m12(t30, 799, t17, 79) - t17;
    a30 = m12(t30, -79, t17, 799) + t30;
    a18 = m12(t29, 79, t18, -799) - t29;
    a29 = m12(t29, 799, t18, 79) - t18;
    a21 = m11(t26, 1703, t21, -1138);
    a26 = m11(t26, 1138, t21, 1703);
    a22 = m11(t25, -1138, t22, -1703);
    a25 = m11(t25, 1703, t22, -1138);

    a16 = cl(t16 + t19);
    t17 = cl(a17 + a18);
    t18 = cl(a17 - a18);
    a19 = cl(t16 - t19);
    a20 = cl(t23 - t20);
    t21 = cl(a22 - a21);
    t22 = cl(a22 + a21);
    a23 = cl(t23 + t20);
    a24 = cl(t24 + t27);
    t25 = cl(a25 + a26);
    t26 = cl(a25 - a26);
    a27 = cl(t24 - t27);
    a28 = cl(t31 - t28);
    t29 = cl(a30 - a29);
    t30 = cl(a30 + a29);
    a31 = cl(t31 + t28);

    a18 = m12(t29, 1567, t18, 312) - t18;
    a29 = m12(t29, -312, t18, 1567) + t29;
    t19 = m12(a28, 1567, a19, 312) - a19;
    t28 = m12(a28, -312, a19, 1567) + a28;
    t20 = m12(a27, 312, a20, -1567) - a27;
    t27 = m12(a27, 1567, a20, 312) - a20;
    a21 = m12(t26, 312, t21, -1567) - t26;
    a26 = m12(t26, 1567, t21, 312) - t21;

    t16 = cl(a16 + a23);
    a17 = cl(t17 + t22);
    t18 = cl(a18 + a21);
    a19 = cl(t19 + t20);
    a20 = cl(t19 - t20);
    t21 = cl(a18 - a21);
    a22 = cl(t17 - t22);
    t23 = cl(a16 - a23);
    t24 = cl(a31 - a24);
    a25 = cl(t30 - t25);
    t26 = cl(a29 - a26);
    a27 = cl(t28 - t27);
    a28 = cl(t28 + t27);
    t29 = cl(a29 + a26);
    a30 = cl(t30 + t25);
    t31 = cl(a31 + a24);

    t20 = s8(a27 - a20);
    t27 = s8(a27 + a20);
    a21 = s8(t26 - t21);
    a26 = s8(t26 + t21);
    t22 = s8(a25 - a22);
    t25 = s8(a25 + a22);
    a23 = s8(t24 - t23);
    a24 = s8(t24 + t23);

    const int o[16] = { t31, a30, t29, a28, t27, a26, t25, a24,
                        a23, t22, a21, t20, a19, t18, a17, t16 };
    int e[16];
#pragma unroll
    for (int k = 0; k < 16; k++) e[k] = c[2 * k * S];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        c[k * S] = cl(e[k] + o[k]);
        c[(31 - k) * S] = cl(e[k] - o[k]);
    }
}

// 64-point: only the first 32 inputs are read (upper half is zero by
// construction of AV1's 64-length transforms), itx_1d.c:436-781.
template <int S> HD void idct64(int *c, const Clamp cl) {
    idct32<2 * S, true>(c, cl);
    const int i1 = c[S], i3 = c[3 * S], i5 = c[5 * S], i7 = c[7 * S];
    const int i9 = c[9 * S], i11 = c[11 * S], i13 = c[13 * S], i15 = c[15 * S];
    const int i17 = c[17 * S], i19 = c[19 * S], i21 = c[21 * S], i23 = c[23 * S];
    const int i25 = c[25 * S], i27 = c[27 * S], i29 = c[29 * S], i31 = c[31 * S];

    int a32 = r12(i1, 101), a33 = r12(i31, -2824), a34 = r12(i17, 1660), a35 = r12(i15, -1474);
    int a36 = r12(i9, 897), a37 = r12(i23, -2191), a38 = r12(i25, 2359), a39 = r12(i7, -700);
    int a40 = r12(i5, 501), a41 = r12(i27, -2520), a42 = r12(i21, 2019), a43 = r12(i11, -1092);
    int a44 = r12(i13, 1285), a45 = r12(i19, -1842), a46 = r12(i29, 2675), a47 = r12(i3, -301);
    int a48 = r12(i3, 4085), a49 = r12(i29, 3102), a50 = r12(i19, 3659), a51 = r12(i13, 3889);
    int a52 = r12(i11, 3948), a53 = r12(i21, 3564), a54 = r12(i27, 3229), a55 = r12(i5, 4065);
    int a56 = r12(i7, 4036), a57 = r12(i25, 3349), a58 = r12(i23, 3461), a59 = r12(i9, 3996);
    int a60 = r12(i15, 3822), a61 = r12(i17, 3745), a62 = r12(i31, 2967), a63 = r12(i1, 4095);

    int t32 = cl(a32 + a33), t33 = cl(a32 - a33), t34 = cl(a35 - a34), t35 = cl(a35 + a34);
    int t36 = cl(a36 + a37), t37 = cl(a36 - a37), t38 = cl(a39 - a38), t39 = cl(a39 + a38);
    int t40 = cl(a40 + a41), t41 = cl(a40 - a41), t42 = cl(a43 - a42), t43 = cl(a43 + a42);
    int t44 = cl(a44 + a45), t45 = cl(a44 - a45), t46 = cl(a47 - a46), t47 = cl(a47 + a46);
    int t48 = cl(a48 + a49), t49 = cl(a48 - a49), t50 = cl(a51 - a50), t51 = cl(a51 + a50);
    int t52 = cl(a52 + a53), t53 = cl(a52 - a53), t54 = cl(a55 - a54), t55 = cl(a55 + a54);
    int t56 = cl(a56 + a57), t57 = cl(a56 - a57), t58 = cl(a59 - a58), t59 = cl(a59 + a58);
    int t60 = cl(a60 + a61), t61 = cl(a60 - a61), t62 = cl(a63 - a62), t63 = cl(a63 + a62);

    a33 = m12(t33, 20, t62, 401) - t33;
    a34 = m12(t34, -401, t61, 20) - t61;
    a37 = m11(t37, -1299, t58, 1583);
    a38 = m11(t38, -1583, t57, -1299);
    a41 = m12(t41, 484, t54, 1931) - t41;
    a42 = m12(t42, -1931, t53, 484) - t53;
    a45 = m12(t45, -1189, t50, -176) + t50;
    a46 = m12(t46, 176, t49, -1189) - t46;
    a49 = m12(t46, -1189, t49, -176) + t49;
    a50 = m12(t45, -176, t50, 1189) + t45;
    a53 = m12(t42, 484, t53, 1931) - t42;
    a54 = m12(t41, 1931, t54, -484) + t54;
    a57 = m11(t38, -1299, t57, 1583);
    a58 = m11(t37, 1583, t58, 1299);
    a61 = m12(t34, 20, t61, 401) - t34;
    a62 = m12(t33, 401, t62, -20) + t62;

    a32 = cl(t32 + t35);
    t33 = cl(a33 + a34);
    t34 = cl(a33 - a34);
    a35 = cl(t32 - t35);
    a36 = cl(t39 - t36);
    t37 = cl(a38 - a37);
    t38 = cl(a38 + a37);
    a39 = cl(t39 + t36);
    a40 = cl(t40 + t43);
    t41 = cl(a41 + a42);
    t42 = cl(a41 - a42);
    a43 = cl(t40 - t43);
    a44 = cl(t47 - t44);
    t45 = cl(a46 - a45);
    t46 = cl(a46 + a45);
    a47 = cl(t47 + t44);
    a48 = cl(t48 + t51);
    t49 = cl(a49 + a50);
    t50 = cl(a49 - a50);
    a51 = cl(t48 - t51);
    a52 = cl(t55 - t52);
    t53 = cl(a54 - a53);
    t54 = cl(a54 + a53);
    a55 = cl(t55 + t52);
    a56 = cl(t56 + t59);
    t57 = cl(a57 + a58);
    t58 = cl(a57 - a58);
    a59 = cl(t56 - t59);
    a60 = cl(t63 - t60);
    t61 = cl(a62 - a61);
    t62 = cl(a62 + a61);
    a63 = cl(t63 + t60);

    a34 = m12(t34, 79, t61, 799) - t34;
    t35 = m12(a35, 79, a60, 799) - a35;
    t36 = m12(a36, -799, a59, 79) - a59;
    a37 = m12(t37, -799, t58, 79) - t58;
    a42 = m11(t42, -1138, t53, 1703);
    t43 = m11(a43, -1138, a52, 1703);
    t44 = m11(a44, -1703, a51, -1138);
    a45 = m11(t45, -1703, t50, -1138);
    a50 = m11(t45, -1138, t50, 1703);
    t51 = m11(a44, -1138, a51, 1703);
    t52 = m11(a43, 1703, a52, 1138);
    a53 = m11(t42, 1703, t53, 1138);
    a58 = m12(t37, 79, t58, 799) - t37;
    t59 = m12(a36, 79, a59, 799) - a36;
    t60 = m12(a35, 799, a60, -79) + a60;
    a61 = m12(t34, 799, t61, -79) + t61;

    t32 = cl(a32 + a39);
    a33 = cl(t33 + t38);
    t34 = cl(a34 + a37);
    a35 = cl(t35 + t36);
    a36 = cl(t35 - t36);
    t37 = cl(a34 - a37);
    a38 = cl(t33 - t38);
    t39 = cl(a32 - a39);
    t40 = cl(a47 - a40);
    a41 = cl(t46 - t41);
    t42 = cl(a45 - a42);
    a43 = cl(t44 - t43);
    a44 = cl(t44 + t43);
    t45 = cl(a45 + a42);
    a46 = cl(t46 + t41);
    t47 = cl(a47 + a40);
    t48 = cl(a48 + a55);
    a49 = cl(t49 + t54);
    t50 = cl(a50 + a53);
    a51 = cl(t51 + t52);
    a52 = cl(t51 - t52);
    t53 = cl(a50 - a53);
    a54 = cl(t49 - t54);
    t55 = cl(a48 - a55);
    t56 = cl(a63 - a56);
    a57 = cl(t62 - t57);
    t58 = cl(a61 - a58);
    a59 = cl(t60 - t59);
    a60 = cl(t60 + t59);
    t61 = cl(a61 + a58);
    a62 = cl(t62 + t57);
    t63 = cl(a63 + a56);

    t36 = m12(a36, 312, a59, 1567) - a36;
    a37 = m12(t37, 312, t58, 1567) - t37;
    t38 = m12(a38, 312, a57, 1567) - a38;
    a39 = m12(t39, 312, t56, 1567) - t39;
    a40 = m12(t40, -1567, t55, 312) - t55;
    t41 = m12(a41, -1567, a54, 312) - a54;
    a42 = m12(t42, -1567, t53, 312) - t53;
    t43 = m12(a43, -1567, a52, 312) - a52;
    t52 = m12(a43, 312, a52, 1567) - a43;
    a53 = m12(t42, 312, t53, 1567) - t42;
    t54 = m12(a41, 312, a54, 1567) - a41;
    a55 = m12(t40, 312, t55, 1567) - t40;
    a56 = m12(t39, 1567, t56, -312) + t56;
    t57 = m12(a38, 1567, a57, -312) + a57;
    a58 = m12(t37, 1567, t58, -312) + t58;
    t59 = m12(a36, 1567, a59, -312) + a59;

    a32 = cl(t32 + t47);
    t33 = cl(a33 + a46);
    a34 = cl(t34 + t45);
    t35 = cl(a35 + a44);
    a36 = cl(t36 + t43);
    t37 = cl(a37 + a42);
    a38 = cl(t38 + t41);
    t39 = cl(a39 + a40);
    t40 = cl(a39 - a40);
    a41 = cl(t38 - t41);
    t42 = cl(a37 - a42);
    a43 = cl(t36 - t43);
    t44 = cl(a35 - a44);
    a45 = cl(t34 - t45);
    t46 = cl(a33 - a46);
    a47 = cl(t32 - t47);
    a48 = cl(t63 - t48);
    t49 = cl(a62 - a49);
    a50 = cl(t61 - t50);
    t51 = cl(a60 - a51);
    a52 = cl(t59 - t52);
    t53 = cl(a58 - a53);
    a54 = cl(t57 - t54);
    t55 = cl(a56 - a55);
    t56 = cl(a56 + a55);
    a57 = cl(t57 + t54);
    t58 = cl(a58 + a53);
    a59 = cl(t59 + t52);
    t60 = cl(a60 + a51);
    a61 = cl(t61 + t50);
    t62 = cl(a62 + a49);
    a63 = cl(t63 + t48);

    a40 = s8(t55 - t40);
    t41 = s8(a54 - a41);
    a42 = s8(t53 - t42);
    t43 = s8(a52 - a43);
    a44 = s8(t51 - t44);
    t45 = s8(a50 - a45);
    a46 = s8(t49 - t46);
    t47 = s8(a48 - a47);
    t48 = s8(a47 + a48);
    a49 = s8(t46 + t49);
    t50 = s8(a45 + a50);
    a51 = s8(t44 + t51);
    t52 = s8(a43 + a52);
    a53 = s8(t42 + t53);
    t54 = s8(a41 + a54);
    a55 = s8(t40 + t55);

    const int o[32] = { a63, t62, a61, t60, a59, t58, a57, t56,
                        a55, t54, a53, t52, a51, t50, a49, t48,
                        t47, a46, t45, a44, t43, a42, t41, a40,
                        t39, a38, t37, a36, t35, a34, t33, a32 };
    int e[32];
#pragma unroll
    for (int k = 0; k < 32; k++) e[k] = c[2 * k * S];
#pragma unroll
    for (int k = 0; k < 32; k++) {
        c[k * S] = cl(e[k] + o[k]);
        c[(63 - k) * S] = cl(e[k] - o[k]);
    }
}

// ---------------------------------------------------------------- ADST
// FLIP writes the outputs in reversed order (itx_1d.c:964-979).
template <int S, bool FLIP> HD void iadst4(int *c, const Clamp) {
    const int i0 = c[0], i1 = c[S], i2 = c[2 * S], i3 = c[3 * S];
    const int o0 = ((1321 * i0 + -293 * i2 + -1614 * i3 + -752 * i1 + 2048) >> 12) + i2 + i3 + i1;
    const int o1 = ((-1614 * i0 - 1321 * i2 - -293 * i3 + -752 * i1 + 2048) >> 12) + i0 - i3 + i1;
    const int o2 = (209 * (i0 - i2 + i3) + 128) >> 8;
    const int o3 = ((-293 * i0 + -1614 * i2 - 1321 * i3 - -752 * i1 + 2048) >> 12) + i0 + i2 - i1;
    c[(FLIP ? 3 : 0) * S] = o0;
    c[(FLIP ? 2 : 1) * S] = o1;
    c[(FLIP ? 1 : 2) * S] = o2;
    c[(FLIP ? 0 : 3) * S] = o3;
}

template <int S, bool FLIP> HD void iadst8(int *c, const Clamp cl) {
    const int i0 = c[0], i1 = c[S], i2 = c[2 * S], i3 = c[3 * S];
    const int i4 = c[4 * S], i5 = c[5 * S], i6 = c[6 * S], i7 = c[7 * S];

    const int a0 = m12(i7, -20, i0, 401) + i7;
    const int a1 = m12(i7, 401, i0, 20) - i0;
    const int a2 = m12(i5, -484, i2, 1931) + i5;
    const int a3 = m12(i5, 1931, i2, 484) - i2;
    int a4 = m11(i3, 1299, i4, 1583);
    int a5 = m11(i3, 1583, i4, -1299);
    int a6 = m12(i1, 1189, i6, -176) + i6;
    int a7 = m12(i1, -176, i6, -1189) + i1;

    const int t0 = cl(a0 + a4), t1 = cl(a1 + a5);
    int t2 = cl(a2 + a6), t3 = cl(a3 + a7);
    const int t4 = cl(a0 - a4), t5 = cl(a1 - a5);
    int t6 = cl(a2 - a6), t7 = cl(a3 - a7);

    a4 = m12(t4, -312, t5, 1567) + t4;
    a5 = m12(t4, 1567, t5, 312) - t5;
    a6 = m12(t7, -312, t6, -1567) + t7;
    a7 = m12(t7, 1567, t6, -312) + t6;

    int o[8];
    o[0] = cl(t0 + t2);
    o[7] = -cl(t1 + t3);
    t2 = cl(t0 - t2);
    t3 = cl(t1 - t3);
    o[1] = -cl(a4 + a6);
    o[6] = cl(a5 + a7);
    t6 = cl(a4 - a6);
    t7 = cl(a5 - a7);
    o[3] = -s8(t2 + t3);
    o[4] = s8(t2 - t3);
    o[2] = s8(t6 + t7);
    o[5] = -s8(t6 - t7);
#pragma unroll
    for (int k = 0; k < 8; k++) c[(FLIP ? 7 - k : k) * S] = o[k];
}

template <int S, bool FLIP> HD void iadst16(int *c, const Clamp cl) {
    int in[16];
#pragma unroll
    for (int k = 0; k < 16; k++) in[k] = c[k * S];

    int t0 = m12(in[15], -5, in[0], 201) + in[15];
    int t1 = m12(in[15], 201, in[0], 5) - in[0];
    int t2 = m12(in[13], -123, in[2], 995) + in[13];
    int t3 = m12(in[13], 995, in[2], 123) - in[2];
    int t4 = m12(in[11], -393, in[4], 1751) + in[11];
    int t5 = m12(in[11], 1751, in[4], 393) - in[4];
    int t6 = m11(in[9], 1645, in[6], 1220);
    int t7 = m11(in[9], 1220, in[6], -1645);
    int t8 = m12(in[7], 2751, in[8], -1061) + in[8];
    int t9 = m12(in[7], -1061, in[8], -2751) + in[7];
    int t10 = m12(in[5], 2106, in[10], -583) + in[10];
    int t11 = m12(in[5], -583, in[10], -2106) + in[5];
    int t12 = m12(in[3], 1380, in[12], -239) + in[12];
    int t13 = m12(in[3], -239, in[12], -1380) + in[3];
    int t14 = m12(in[1], 601, in[14], -44) + in[14];
    int t15 = m12(in[1], -44, in[14], -601) + in[1];

    int a0 = cl(t0 + t8), a1 = cl(t1 + t9), a2 = cl(t2 + t10), a3 = cl(t3 + t11);
    int a4 = cl(t4 + t12), a5 = cl(t5 + t13), a6 = cl(t6 + t14), a7 = cl(t7 + t15);
    int a8 = cl(t0 - t8), a9 = cl(t1 - t9), a10 = cl(t2 - t10), a11 = cl(t3 - t11);
    int a12 = cl(t4 - t12), a13 = cl(t5 - t13), a14 = cl(t6 - t14), a15 = cl(t7 - t15);

    t8 = m12(a8, -79, a9, 799) + a8;
    t9 = m12(a8, 799, a9, 79) - a9;
    t10 = m12(a10, 2276, a11, -690) + a11;
    t11 = m12(a10, -690, a11, -2276) + a10;
    t12 = m12(a13, -79, a12, -799) + a13;
    t13 = m12(a13, 799, a12, -79) + a12;
    t14 = m12(a15, 2276, a14, 690) - a14;
    t15 = m12(a15, -690, a14, 2276) + a15;

    t0 = cl(a0 + a4);
    t1 = cl(a1 + a5);
    t2 = cl(a2 + a6);
    t3 = cl(a3 + a7);
    t4 = cl(a0 - a4);
    t5 = cl(a1 - a5);
    t6 = cl(a2 - a6);
    t7 = cl(a3 - a7);
    a8 = cl(t8 + t12);
    a9 = cl(t9 + t13);
    a10 = cl(t10 + t14);
    a11 = cl(t11 + t15);
    a12 = cl(t8 - t12);
    a13 = cl(t9 - t13);
    a14 = cl(t10 - t14);
    a15 = cl(t11 - t15);

    a4 = m12(t4, -312, t5, 1567) + t4;
    a5 = m12(t4, 1567, t5, 312) - t5;
    a6 = m12(t7, -312, t6, -1567) + t7;
    a7 = m12(t7, 1567, t6, -312) + t6;
    t12 = m12(a12, -312, a13, 1567) + a12;
    t13 = m12(a12, 1567, a13, 312) - a13;
    t14 = m12(a15, -312, a14, -1567) + a15;
    t15 = m12(a15, 1567, a14, -312) + a14;

    int o[16];
    o[0] = cl(t0 + t2);
    o[15] = -cl(t1 + t3);
    a2 = cl(t0 - t2);
    a3 = cl(t1 - t3);
    o[3] = -cl(a4 + a6);
    o[12] = cl(a5 + a7);
    t6 = cl(a4 - a6);
    t7 = cl(a5 - a7);
    o[1] = -cl(a8 + a10);
    o[14] = cl(a9 + a11);
    t10 = cl(a8 - a10);
    t11 = cl(a9 - a11);
    o[2] = cl(t12 + t14);
    o[13] = -cl(t13 + t15);
    a14 = cl(t12 - t14);
    a15 = cl(t13 - t15);

    o[7] = -s8(a2 + a3);
    o[8] = s8(a2 - a3);
    o[4] = s8(t6 + t7);
    o[11] = -s8(t6 - t7);
    o[6] = s8(t10 + t11);
    o[9] = -s8(t10 - t11);
    o[5] = -s8(a14 + a15);
    o[10] = s8(a14 - a15);
#pragma unroll
    for (int k = 0; k < 16; k++) c[(FLIP ? 15 - k : k) * S] = o[k];
}

// ---------------------------------------------------------------- identity
template <int S, int N> HD void iidentity(int *c) {
#pragma unroll
    for (int k = 0; k < N; k++) {
        const int v = c[k * S];
        if (N == 4) c[k * S] = v + ((v * 1697 + 2048) >> 12);
        else if (N == 8) c[k * S] = v * 2;
        else if (N == 16) c[k * S] = 2 * v + ((v * 1697 + 1024) >> 11);
        else c[k * S] = v * 4;
    }
}

// ---------------------------------------------------------------- WHT (lossless)
template <int S> HD void iwht4(int *c) {
    const int i0 = c[0], i1 = c[S], i2 = c[2 * S], i3 = c[3 * S];
    const int t0 = i0 + i1;
    const int t2 = i2 - i3;
    const int t4 = (t0 - t2) >> 1;
    const int t3 = t4 - i3;
    const int t1 = t4 - i1;
    c[0] = t0 - t3;
    c[S] = t3;
    c[2 * S] = t1;
    c[3 * S] = t2 + t1;
}

// 1-D kernel selector used by the 2-D drivers.
enum Itx1d { K_DCT = 0, K_ADST = 1, K_FLIPADST = 2, K_IDENTITY = 3, K_WHT = 4 };

// Run the N-point transform `kind` on c[0..N-1] (stride 1).  For N == 64 only
// c[0..31] are inputs.  Returns false for combinations AV1 does not define.
template <int N> HD void itx1d_run(int *c, const int kind, const Clamp cl) {
    if (N == 4) {
        switch (kind) {
        case K_DCT: idct4<1, false>(c, cl); break;
        case K_ADST: iadst4<1, false>(c, cl); break;
        case K_FLIPADST: iadst4<1, true>(c, cl); break;
        case K_IDENTITY: iidentity<1, 4>(c); break;
        default: iwht4<1>(c); break;
        }
    } else if (N == 8) {
        switch (kind) {
        case K_DCT: idct8<1, false>(c, cl); break;
        case K_ADST: iadst8<1, false>(c, cl); break;
        case K_FLIPADST: iadst8<1, true>(c, cl); break;
        default: iidentity<1, 8>(c); break;
        }
    } else if (N == 16) {
        switch (kind) {
        case K_DCT: idct16<1, false>(c, cl); break;
        case K_ADST: iadst16<1, false>(c, cl); break;
        case K_FLIPADST: iadst16<1, true>(c, cl); break;
        default: iidentity<1, 16>(c); break;
        }
    } else if (N == 32) {
        if (kind == K_DCT) idct32<1, false>(c, cl);
        else iidentity<1, 32>(c);
    } else {
        idct64<1>(c, cl);
    }
}

// Same with the knowledge that inputs c[NZ..] are zero: writing them as literal
// zeros lets the compiler fold the butterflies that only see zeros (exact:
// every stage maps 0 -> 0).  `nz` = number of leading inputs that may be
// non-zero; buckets of 8 / 16 / all.
template <int N, int NZ> HD void itx1d_run_nz(int *c, const int kind, const Clamp cl) {
    constexpr int NIN = N == 64 ? 32 : N;
#pragma unroll
    for (int k = NZ; k < NIN; k++) c[k] = 0;
    itx1d_run<N>(c, kind, cl);
}
template <int N> HD void itx1d_dispatch(int *c, const int kind, const Clamp cl, const int nz) {
    constexpr int NIN = N == 64 ? 32 : N;
    if (NIN > 8 && nz <= 8) itx1d_run_nz<N, (NIN > 8 ? 8 : NIN)>(c, kind, cl);
    else if (NIN > 16 && nz <= 16) itx1d_run_nz<N, (NIN > 16 ? 16 : NIN)>(c, kind, cl);
    else itx1d_run<N>(c, kind, cl);
}

}  // namespace d1
