// Exact mode of the device libm (MARAY_LIBM=glibc): sin / exp / log that return glibc's bits.
//
// The reference evaluates Expr::Sin/Exp/Ln with the platform libm (src/lib.rs:648-650; the WASM path imports the
// same functions, src/wasm.rs:11-13).  On the hosts this project runs on that is glibc 2.39, whose x86-64 build
// selects the FMA variants (__sin_fma, __exp_fma, __log_fma) on any CPU with AVX2+FMA.  None of the three is
// correctly rounded (0.55 / 0.51 / 0.52 ULP), so "a correctly rounded device libm" would still differ from the
// reference in the last bit of ~1 % of arguments; the only way to the same bits is the same arithmetic.  This file
// restates those three routines operation for operation -- the same tables (csrc/glibc_libm_tables.inc, read from
// the host's libm.so.6 by tools/extract_glibc_libm.py), the same polynomial shapes and, because the library is
// compiled with -mfma and contraction on, the same FUSED operations, which were read off the machine code of the
// shipped library (every MR_FMA below is a vfmadd/vfnmadd/vfmsub there, every plain * + - a vmulsd/vaddsd/vsubsd).
// Algorithms: exp and log are Szabolcs Nagy's (ARM optimized routines: sysdeps/ieee754/dbl-64/e_exp.c, e_log.c),
// sin is the IBM Accurate Mathematical Library's table method (sysdeps/ieee754/dbl-64/s_sin.c: do_sin, do_cos,
// reduce_sincos, TAYLOR_SIN).
//
// Pinned by tools/glibc_libm_check.c (tests/test_glibc_libm.py): the host rendition of this text (MR_LIBM_HOST)
// against the host's sin/exp/log, bit for bit, >= 10^7 arguments per range plus the special values.
//
// One documented gap: |x| >= 105414350 in sin.  glibc reduces those with __branred (a Payne-Hanek reduction over a
// 1 KB table of 2/pi digits); here they go to libdevice's sin (<= 2 ULP).  NaN payloads are not preserved either
// (every NaN result is the canonical quiet NaN; `as u8` maps all of them to 0).
#ifndef MARAY_DEVICE_LIBM_GLIBC_CUH
#define MARAY_DEVICE_LIBM_GLIBC_CUH

#ifdef MR_LIBM_HOST
#define MRG_TABLE_U64 static const unsigned long long
#define MRG_TABLE_F64 static const double
#define MRG_COLD static double
#define MRG_LDG2U(p, a, b) do { (a) = (p)[0]; (b) = (p)[1]; } while (0)
#define MRG_HUGE_SIN(x) sin(x)
#else
#define MRG_TABLE_U64 static __device__ __align__(16) const unsigned long long
#define MRG_TABLE_F64 static __device__ __align__(32) const double
#define MRG_COLD static __device__ __noinline__ double
#define MRG_LDG2U(p, a, b) do { ulonglong2 v2_ = __ldg(reinterpret_cast<const ulonglong2*>(p)); (a) = v2_.x; (b) = v2_.y; } while (0)
#define MRG_HUGE_SIN(x) sin(x)
#endif
#define K MRG_SIN_K
#include "glibc_libm_tables.inc"

MR_TABLE double MRG_EXP_K[8] = {MRG_EXP_CONSTS};    // invln2N, shift, negln2hiN, negln2loN, C2, C3, C4, C5
MR_TABLE double MRG_LOG_K[18] = {MRG_LOG_CONSTS};   // ln2hi, ln2lo, A[0..4], B[0..10]
// s_sin.c / usncs.h / trigo.h constants
MR_TABLE double MRG_SIN_K[] = {
    /* 0 */ 0x1.8p+45,                    // big: ulp = 2^-7, the table step
    /* 1 */ -0x1.5555555555515p-3,        // sn3
    /* 2 */ 0x1.11110e829872fp-7,         // sn5
    /* 3 */ -0x1.5555555555535p-5,        // cs4   (cs2 = 0.5)
    /* 4 */ 0x1.6c16bedd9e239p-10,        // cs6
    /* 5 */ -0x1.5555555555555p-3,        // s1
    /* 6 */ 0x1.1111111110ecep-7,         // s2
    /* 7 */ -0x1.a01a019db08b8p-13,       // s3
    /* 8 */ 0x1.71de27b9a7ed9p-19,        // s4
    /* 9 */ -0x1.addffc2fcdf59p-26,       // s5
    /* 10 */ 0x1.921fb54442d18p+0,        // hp0
    /* 11 */ 0x1.1a62633145c07p-54,       // hp1
    /* 12 */ 0x1.45f306dc9c883p-1,        // hpinv
    /* 13 */ 0x1.8p+52,                   // toint
    /* 14 */ 0x1.921fb58000000p+0,        // mp1
    /* 15 */ -0x1.dde973c000000p-27,      // mp2
    /* 16 */ -0x1.cb3b398000000p-55,      // pp3
    /* 17 */ -0x1.d747f23e32ed7p-83,      // pp4
};

MR_FN double mrg_flip(double v, unsigned int sign_bit) {   // v with its sign bit xor-ed (integer pipe)
    return mr_hilo((int)((unsigned int)mr_hi32(v) ^ sign_bit), mr_lo32(v));
}
MR_FN double mrg_abs(double v) { return mr_hilo(mr_hi32(v) & 0x7fffffff, mr_lo32(v)); }

// ---------------------------------------------------------------------------------------------- exp (e_exp.c)
// Fast range: 2^-54 <= |x| < 512 (abstop - 0x3c9 < 0x3f): no special case can occur.
MR_FN int mr_exp_inrange_g(double x) { return (unsigned int)(((mr_hi32(x) >> 20) & 0x7ff) - 0x3c9) < 0x3fu; }

// The part shared with the special cases: exp(x) = scale * (1 + tmp), scale = 2^(k/128) from the table.
MR_FN double mrg_exp_core(double x, unsigned int* ki_lo, unsigned long long* sbits) {
    const double kd0 = MR_FMA(x, MRG_EXP_K[0], MRG_EXP_K[1]);     // z + Shift, fused
    const double kd = kd0 - MRG_EXP_K[1];
    double r = MR_FMA(kd, MRG_EXP_K[2], x);
    r = MR_FMA(kd, MRG_EXP_K[3], r);
    const unsigned int ki = (unsigned int)mr_lo32(kd0);
    unsigned long long tailb, sb;
    MRG_LDG2U(MRG_EXP_TAB + 2 * (ki & 127u), tailb, sb);
    const double tail = mr_hilo((int)(tailb >> 32), (int)(unsigned int)tailb);
    *sbits = sb + ((unsigned long long)ki << 45);
    *ki_lo = ki;
    const double p23 = MR_FMA(r, MRG_EXP_K[5], MRG_EXP_K[4]);     // C2 + r*C3
    const double tr = r + tail;
    const double r2 = r * r;
    const double p45 = MR_FMA(r, MRG_EXP_K[7], MRG_EXP_K[6]);     // C4 + r*C5
    const double t1 = MR_FMA(p23, r2, tr);
    const double r4 = r2 * r2;
    return MR_FMA(r4, p45, t1);
}
MR_FN double mr_exp_fast_g(double x) {
    unsigned int ki; unsigned long long sbits;
    const double tmp = mrg_exp_core(x, &ki, &sbits);
    const double scale = mr_hilo((int)(sbits >> 32), (int)(unsigned int)sbits);
    return MR_FMA(scale, tmp, scale);
}
// Everything else: tiny, |x| >= 512 (result may overflow, underflow or be subnormal), infinities, NaN.
MRG_COLD mr_exp_slow_g(double x) {
    const unsigned int hi = (unsigned int)mr_hi32(x);
    const unsigned int abstop = (hi >> 20) & 0x7ff;
    if (abstop < 0x3c9u) return 1.0 + x;
    if (abstop >= 0x409u) {
        if (hi == 0xfff00000u && mr_lo32(x) == 0) return 0.0;
        if (abstop >= 0x7ffu) return 1.0 + x;
        return (hi >> 31) ? 0.0 : mr_hilo(0x7ff00000, 0);          // __math_uflow(0) / __math_oflow(0)
    }
    if (abstop < 0x408u) return mr_exp_fast_g(x);                  // not special after all
    unsigned int ki; unsigned long long sbits;
    const double tmp = mrg_exp_core(x, &ki, &sbits);
    if ((ki & 0x80000000u) == 0) {                                 // k > 0: the exponent of scale may have overflowed
        sbits -= 1009ull << 52;
        const double scale = mr_hilo((int)(sbits >> 32), (int)(unsigned int)sbits);
        return 0x1p1009 * MR_FMA(scale, tmp, scale);
    }
    sbits += 1022ull << 52;                                        // k < 0: care in the subnormal range
    const double scale = mr_hilo((int)(sbits >> 32), (int)(unsigned int)sbits);
    const double st = tmp * scale;                                 // not fused in the library
    double y = scale + st;
    if (y < 1.0) {
        double lo = (scale - y) + st;
        const double hi1 = 1.0 + y;
        lo = ((1.0 - hi1) + y) + lo;
        y = (hi1 + lo) - 1.0;
        if (y == 0.0) y = 0.0;
    }
    return 0x1p-1022 * y;
}

// ---------------------------------------------------------------------------------------------- log (e_log.c)
// Fast range: positive, normal, finite.
MR_FN int mr_log_inrange_g(double x) { return (unsigned int)(((unsigned int)mr_hi32(x) >> 16) - 0x10u) < 0x7fe0u; }

MR_FN double mrg_log_main(int hi, int lo) {       // x = 2^k z, z in [OFF, 2 OFF), OFF = 0x3fe6000000000000
    const int th = hi - 0x3fe60000;
    const int i = (th >> 13) & 127;
    const int k = th >> 20;
    const double z = mr_hilo(hi - (int)((unsigned int)th & 0xfff00000u), lo);
    double invc, logc;
    MR_LDG2(MRG_LOG_TAB + 2 * i, invc, logc);
    const double kd = (double)k;
    const double w = MR_FMA(kd, MRG_LOG_K[0], logc);
    const double r = MR_FMA(z, invc, -1.0);
    const double a12 = MR_FMA(r, MRG_LOG_K[4], MRG_LOG_K[3]);      // A[1] + r*A[2]
    const double hi_ = r + w;
    const double r2 = r * r;
    double lo_ = (w - hi_) + r;
    lo_ = MR_FMA(kd, MRG_LOG_K[1], lo_);
    const double r3 = r * r2;
    const double a34 = MR_FMA(r, MRG_LOG_K[6], MRG_LOG_K[5]);      // A[3] + r*A[4]
    lo_ = MR_FMA(r2, MRG_LOG_K[2], lo_);                           // lo + r2*A[0]
    const double p = MR_FMA(a34, r2, a12);
    return MR_FMA(r3, p, lo_) + hi_;
}
#define B(i) MRG_LOG_K[7 + (i)]
MR_FN double mrg_log_near1(double x) {            // 1 - 2^-4 <= x < 1 + 0x1.09p-4, x != 1
    const double r = x - 1.0;
    const double q12 = MR_FMA(r, B(2), B(1));
    const double q45 = MR_FMA(r, B(5), B(4));
    const double r2 = r * r;
    const double q78 = MR_FMA(r, B(8), B(7));
    const double q123 = MR_FMA(r2, B(3), q12);
    const double q456 = MR_FMA(r2, B(6), q45);
    const double r3 = r * r2;
    double q = MR_FMA(r2, B(9), q78);
    q = MR_FMA(r3, B(10), q);
    q = MR_FMA(q, r3, q456);
    q = MR_FMA(q, r3, q123);
    const double t = MR_FMA(r, 0x1p27, r);
    const double rhi = MR_FMA(-0x1p27, r, t);
    const double rhi2 = rhi * rhi;
    const double rlo = r - rhi;
    const double hi = MR_FMA(rhi2, B(0), r);
    const double d = r - hi;
    const double rs = r + rhi;
    double lo = MR_FMA(rhi2, B(0), d);
    lo = MR_FMA(B(0) * rlo, rs, lo);
    return hi + MR_FMA(q, r3, lo);
}
#undef B
MR_FN double mr_log_fast_g(double x) {
    const int hi = mr_hi32(x), lo = mr_lo32(x);
    if ((unsigned int)(hi - 0x3fee0000) < 0x30900u) {
        if (hi == 0x3ff00000 && lo == 0) return 0.0;
        return mrg_log_near1(x);
    }
    return mrg_log_main(hi, lo);
}
MRG_COLD mr_log_slow_g(double x) {
    const unsigned int hi = (unsigned int)mr_hi32(x);
    if (((hi & 0x7fffffffu) | (unsigned int)(mr_lo32(x) != 0)) == 0) return mr_hilo((int)0xfff00000u, 0);   // log(+-0) = -inf
    if (hi == 0x7ff00000u && mr_lo32(x) == 0) return x;                                                     // log(inf) = inf
    if ((hi >> 31) || (hi & 0x7ff00000u) == 0x7ff00000u) return mr_hilo(0x7ff80000, 0);                     // negative, NaN
    if ((hi >> 20) != 0) return mr_log_fast_g(x);                                                           // not special after all
    const double xs = x * 0x1p52;                                                                           // subnormal: normalise
    return mrg_log_main(mr_hi32(xs) - (52 << 20), mr_lo32(xs));
}

// ---------------------------------------------------------------------------------------------- sin (s_sin.c)
// Fast range: 2^-26 <= |x| < 105414350.
MR_FN int mr_sin_inrange_g(double x) { return (unsigned int)((mr_hi32(x) & 0x7fffffff) - 0x3e500000) < (0x419921fbu - 0x3e500000u); }

// do_sin(a, da): sin(a + da) for |a| <= 0.855469, through the table of sin/cos at multiples of 1/128.
MR_FN double mrg_do_sin(double a, double da) {
    const double aa = mrg_abs(a);
    if (aa < 0.126) {                                              // TAYLOR_SIN(a*a, a, da)
        const double xx = a * a;
        double p = MR_FMA(K[9], xx, K[8]);
        p = MR_FMA(p, xx, K[7]);
        p = MR_FMA(p, xx, K[6]);
        p = MR_FMA(p, xx, K[5]);
        const double hd = da * 0.5;
        const double t = MR_FMA(xx, MR_FMA(p, a, -hd), da);
        return t + a;
    }
    const double dx = (a <= 0.0) ? -da : da;
    const double u = aa + K[0];
    const double x = aa - (u - K[0]);
    const double* e = MRG_SINCOS_TAB + 4 * (mr_lo32(u) & 127);   // in range: < 110; the mask keeps any argument in bounds
    double sn, ssn, cs, ccs;
    MR_LDG2(e, sn, ssn);
    MR_LDG2(e + 2, cs, ccs);
    const double xx = x * x;
    const double ps = MR_FMA(xx, K[2], K[1]);                      // sn3 + xx*sn5
    const double s = x + MR_FMA(x * xx, ps, dx);
    double pc = MR_FMA(xx, K[4], K[3]);                            // cs4 + xx*cs6
    pc = MR_FMA(pc, xx, 0.5);
    const double c = MR_FMA(x, dx, xx * pc);
    double cor = MR_FMA(s, ccs, ssn);
    cor = MR_FMA(-c, sn, cor);
    cor = MR_FMA(s, cs, cor);
    const double res = sn + cor;
    return mr_hilo((mr_hi32(res) & 0x7fffffff) | (mr_hi32(a) & (int)0x80000000u), mr_lo32(res));   // copysign(res, a)
}
// do_cos(a, da): cos(a + da), same table.
MR_FN double mrg_do_cos(double a, double da) {
    const double aa = mrg_abs(a);
    const double dx = (a < 0.0) ? -da : da;
    const double u = aa + K[0];
    const double x = (aa - (u - K[0])) + dx;
    const double* e = MRG_SINCOS_TAB + 4 * (mr_lo32(u) & 127);   // in range: < 110; the mask keeps any argument in bounds
    double sn, ssn, cs, ccs;
    MR_LDG2(e, sn, ssn);
    MR_LDG2(e + 2, cs, ccs);
    const double xx = x * x;
    const double ps = MR_FMA(xx, K[2], K[1]);
    const double s = MR_FMA(x * xx, ps, x);
    double pc = MR_FMA(xx, K[4], K[3]);
    pc = MR_FMA(pc, xx, 0.5);
    const double c = xx * pc;
    double cor = MR_FMA(-s, ssn, ccs);
    cor = MR_FMA(-c, cs, cor);
    cor = MR_FMA(-s, sn, cor);
    return cs + cor;
}
MR_FN double mr_sin_fast_g(double x) {
    const unsigned int k = (unsigned int)mr_hi32(x) & 0x7fffffffu;
    if (k < 0x3feb6000u) return mrg_do_sin(x, 0.0);               // |x| < 0.855469
    if (k < 0x400368fdu) {                                         // |x| < 2.426265: sin(x) = cos(pi/2 - |x|), sign of x
        const double t = K[10] - mrg_abs(x);
        const double c = mrg_do_cos(t, K[11]);
        return mr_hilo((mr_hi32(c) & 0x7fffffff) | (mr_hi32(x) & (int)0x80000000u), mr_lo32(c));   // copysign(c, x)
    }
    // reduce_sincos: x = xn * pi/2 + (a + da), pi/2 in four pieces (136 bits)
    const double t = MR_FMA(x, K[12], K[13]);
    const double xn = t - K[13];
    double y = MR_FMA(-xn, K[14], x);
    y = MR_FMA(-xn, K[15], y);
    const int n = mr_lo32(t);
    const double t2 = MR_FMA(-xn, K[16], y);
    const double db = MR_FMA(-K[16], xn, y - t2);
    const double a = MR_FMA(-xn, K[17], t2);
    const double da = db + MR_FMA(-xn, K[17], t2 - a);
    const double r = (n & 1) ? mrg_do_cos(a, da) : mrg_do_sin(a, da);
    return mrg_flip(r, ((unsigned int)n & 2u) << 30);
}
MRG_COLD mr_sin_slow_g(double x) {
    const unsigned int k = (unsigned int)mr_hi32(x) & 0x7fffffffu;
    if (k < 0x3e500000u) return x;                                 // |x| < 2^-26
    if (k < 0x419921fbu) return mr_sin_fast_g(x);                  // not special after all
    if (k >= 0x7ff00000u) return mr_hilo(0x7ff80000, 0);           // x / x
    return MRG_HUGE_SIN(x);                                        // the documented gap: glibc's __branred range
}

#undef K
#endif  // MARAY_DEVICE_LIBM_GLIBC_CUH
