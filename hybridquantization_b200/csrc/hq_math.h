// hq_math.h — single-source (host + device) arithmetic of the hot path.
//
// Every function here is built ONLY from IEEE-754 correctly-rounded primitives
// (+, -, *, /, fma, rint, int<->float conversions) with explicit rounding, so the
// same inputs give the same bits on the host compiler (g++ -ffp-contract=off) and
// on sm_100a (explicit __f*_rn / __d*_rn intrinsics: no FMA contraction, no
// approximate division).  No libm transcendental is called: the two that the
// reference needs, (float)Math.pow(b, 2.4f) and (float)Math.pow(t, 1.0/3.0)
// (ScielabProcessor.java:282-284, :301-305), are restated as hq_pow_2p4f and
// hq_cbrtf below and are verified EXHAUSTIVELY against glibc pow() over their whole
// float domains by tests/test_math_exhaustive.py (0 mismatches on 33.5 M and
// 59.7 M floats respectively).
//
// Formulas follow the reference's Java CPU helpers:
//   sRGB decode + RGB->Opp      ScielabProcessor.java:279-291
//   Opp->XYZ->Lab               ScielabProcessor.java:293-311
//   white points, Lab constants ScielabProcessor.java:20-21, :59-61
// Distance / tie rule follow OptimizedConvolution.cl:178-193 (see hq_dist2).
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define HQ_HD __host__ __device__ __forceinline__
#else
#define HQ_HD static inline
#endif

// ---------------------------------------------------------------- IEEE primitives
#if defined(__CUDA_ARCH__)
#define HQ_FADD(a, b) __fadd_rn((a), (b))
#define HQ_FSUB(a, b) __fsub_rn((a), (b))
#define HQ_FMUL(a, b) __fmul_rn((a), (b))
#define HQ_FDIV(a, b) __fdiv_rn((a), (b))
#define HQ_FFMA(a, b, c) __fmaf_rn((a), (b), (c))
#define HQ_FSQRT(a) __fsqrt_rn((a))
#define HQ_DADD(a, b) __dadd_rn((a), (b))
#define HQ_DSUB(a, b) __dsub_rn((a), (b))
#define HQ_DMUL(a, b) __dmul_rn((a), (b))
#define HQ_DDIV(a, b) __ddiv_rn((a), (b))
#define HQ_DFMA(a, b, c) __fma_rn((a), (b), (c))
#define HQ_DRINT(a) rint((a))
#define HQ_F2U(f) __float_as_uint((f))
#define HQ_U2F(u) __uint_as_float((u))
#define HQ_D2U(d) ((uint64_t)__double_as_longlong((d)))
#define HQ_U2D(u) __longlong_as_double((long long)(u))
#define HQ_F2LL_RN(f) __float2ll_rn((f))
#else
// host: the translation unit MUST be compiled with -ffp-contract=off
#define HQ_FADD(a, b) ((float)(a) + (float)(b))
#define HQ_FSUB(a, b) ((float)(a) - (float)(b))
#define HQ_FMUL(a, b) ((float)(a) * (float)(b))
#define HQ_FDIV(a, b) ((float)(a) / (float)(b))
#define HQ_FFMA(a, b, c) fmaf((a), (b), (c))
#define HQ_FSQRT(a) sqrtf((a))
#define HQ_DADD(a, b) ((double)(a) + (double)(b))
#define HQ_DSUB(a, b) ((double)(a) - (double)(b))
#define HQ_DMUL(a, b) ((double)(a) * (double)(b))
#define HQ_DDIV(a, b) ((double)(a) / (double)(b))
#define HQ_DFMA(a, b, c) fma((a), (b), (c))
#define HQ_DRINT(a) rint((a))
static inline uint32_t hq_f2u_host(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float hq_u2f_host(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint64_t hq_d2u_host(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline double hq_u2d_host(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
#define HQ_F2U(f) hq_f2u_host((f))
#define HQ_U2F(u) hq_u2f_host((u))
#define HQ_D2U(d) hq_d2u_host((d))
#define HQ_U2D(u) hq_u2d_host((u))
#define HQ_F2LL_RN(f) ((long long)llrintf((f)))
#endif

// ---------------------------------------------------------------- constants
// Fixed-point scale of every reduced quantity (error, Lab sums): 2^24.
#define HQ_FX_SCALE 16777216.0f
#define HQ_FX_SCALE_D 16777216.0

// ScielabProcessor.java:59-61, evaluated in fp32 exactly as Java does:
//   LABDELTA = 6f/29f, LABDELTA2 = LABDELTA*LABDELTA, LABDELTA3 = LABDELTA2*LABDELTA
// (bit patterns pinned by tests/test_oracle_golden.py against a run-time evaluation)
#define HQ_LABDELTA3 0x1.22354ep-7f        /* 0.008856452070 */
#define HQ_3LABDELTA2 0x1.070050p-3f       /* 3*LABDELTA2 = 0.128418565 */
#define HQ_4_OVER_29 0x1.1a7b96p-3f        /* 4f/29f = 0.137931034 */
#define HQ_RCP_3LABDELTA2 0x1.f25eccp+2f    /* RN(1 / (3*LABDELTA2)) = 7.787036 */

enum { HQ_WHITE_D65 = 0, HQ_WHITE_D50 = 1 };

struct hq_float3 {
    float x, y, z;
};

// ScielabProcessor.java:20-21
HQ_HD hq_float3 hq_whitepoint(int wp) {
    hq_float3 w;
    if (wp == HQ_WHITE_D50) { w.x = 0.966797f; w.y = 1.0f; w.z = 0.825188f; }
    else { w.x = 0.95047f; w.y = 1.0f; w.z = 1.0883f; }
    return w;
}

// exact float -> double of a positive NORMAL float with integer operations (on sm_100a the
// F2F conversion runs on the 16-lane XU pipe, which was the busiest unit of rgb_to_lab_kernel)
HQ_HD double hq_widen_pos(float f) {
#if defined(__CUDA_ARCH__)
    const uint32_t b = __float_as_uint(f);
    return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
#else
    return (double)f;
#endif
}

// ---------------------------------------------------------------- (float)pow(t, 1.0/3.0)
// Restates `(float)Math.pow(t, 1.0 / 3.0)` (ScielabProcessor.java:301,303,305) for
// positive normal t.  REFERENCE ROUTINE (host, and the device's fallback): inverse-cube-root
// Newton in fp32 from a seed, then two Newton steps on y^3 = t in fp64 whose residual is formed
// with one fma; the fp64 value is within 1 ulp(double) of the true root before the final narrowing.
HQ_HD float hq_cbrtf_f64(float t) {
#if defined(__CUDA_ARCH__)
    // device seed: x ~ t^(-1/3) = 2^(-log2(t)/3) from the SFU approximations (rel. error < 1e-6).
    // The SFU results are not reproducible on the host, but the value returned below does not
    // depend on the seed unless the true root sits within ~1e-17 of a rounding boundary; that
    // this never happens is verified EXHAUSTIVELY on the B200 against the oracle for every float
    // of the domain (tests/test_gpu_parity.py::test_device_math_exhaustive).
    float lg, x;
    asm("lg2.approx.f32 %0, %1;" : "=f"(lg) : "f"(t));
    asm("ex2.approx.f32 %0, %1;" : "=f"(x) : "f"(HQ_FMUL(lg, -0x1.555556p-2f)));
#else
    const uint32_t bits = HQ_F2U(t);
    float x = HQ_U2F(0x54a2fa8cu - bits / 3u);  // x ~ t^(-1/3), |rel err| < 3.3 %
    const float t3 = HQ_FMUL(t, -0x1.555556p-2f);  // -t/3
    const float c43 = 0x1.555556p+0f;              // 4/3
    for (int i = 0; i < 3; ++i) {  // x <- x * (4/3 - t/3 * x^3): e -> -2 e^2
        const float x2 = HQ_FMUL(x, x);
        const float x3 = HQ_FMUL(x2, x);
        const float w = HQ_FFMA(t3, x3, c43);
        x = HQ_FMUL(x, w);
    }
#endif
    const double xd = hq_widen_pos(x), td = hq_widen_pos(t);
    const double x2 = HQ_DMUL(xd, xd);
    double y = HQ_DMUL(td, x2);                         // t * t^(-2/3)
    const double h = HQ_DMUL(x2, 0x1.5555555555555p-2);  // ~ 1 / (3 y^2)
#pragma unroll
    for (int i = 0; i < 2; ++i) {  // Newton on y^3 = t, residual with one fma: e -> 16 eps^2 -> ~1e-18
        const double s = HQ_DMUL(y, y);
        const double r = HQ_DFMA(s, y, -td);
        y = HQ_DFMA(-r, h, y);
    }
    return (float)y;
}

#if defined(__CUDACC__)
// DEVICE FAST PATH, fp32 only.  y0 = 2^(log2(t)/3) from the SFU (rel. error ~2^-21), then ONE Newton
// correction whose residual t - y0^3 is formed without cancellation error: y0*y0 = q + ql and
// q*y0 = p + pl exactly (fma), t - p is exact (Sterbenz), so r = ((t - p) - pl) - ql*y0 carries a
// relative error of ~2^-24 only.  d = r / (3 q) is the correction (|d| <= ~8 ulp, known to ~2^-18 ulp
// including Newton's second-order term).  The root is y0 + d: it is rounded twice, with d moved down and up
// by y0 * 2^-39 (2^-16 .. 2^-15 ulp, several times the uncertainty of d).  When both roundings agree that
// float is the correctly rounded root; when they differ (the root is within 2^-15 ulp of a rounding
// boundary, one input in ~16,000) the fp64 routine above decides.  The boundaries of (float)pow(t, 1.0/3.0)
// (double pow, then narrowed) sit at the same midpoints to 2^-29 ulp, far inside that window.  Equality with
// the oracle for EVERY float of the domain is verified on the B200 (test_device_math_exhaustive).
// The SFU operations carry .ftz: all operands are normal floats on the domain, .ftz only removes the
// denormal pre-scaling code nvcc otherwise wraps around MUFU.
#define HQ_CBRT_ETA 0x1p-39f
// the rare fp64 decision as a CALL: inlined, its ~60 instructions sat in every unrolled copy of every caller's epilogue
static __device__ __noinline__ float hq_cbrtf_f64_call(float t) { return hq_cbrtf_f64(t); }
__device__ __forceinline__ float hq_cbrtf_fast(float t) {
    float lg, y0, rq;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(t));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(__fmul_rn(lg, 0x1.555556p-2f)));
    const float q = __fmul_rn(y0, y0);
    const float nql = __fmaf_rn(-y0, y0, q);   // q - y0*y0, exact
    const float p = __fmul_rn(q, y0);
    const float npl = __fmaf_rn(q, -y0, p);    // p - q*y0, exact
    const float r = __fmaf_rn(nql, y0, __fadd_rn(__fsub_rn(t, p), npl));  // t - y0^3
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rq) : "f"(q));
    const float d = __fmul_rn(__fmul_rn(r, 0x1.555556p-2f), rq);
    const float s_lo = __fadd_rn(y0, __fmaf_rn(y0, -HQ_CBRT_ETA, d));
    const float s_hi = __fadd_rn(y0, __fmaf_rn(y0, HQ_CBRT_ETA, d));
    if (s_lo != s_hi) return hq_cbrtf_f64_call(t);
    return s_lo;
}
#endif
HQ_HD float hq_cbrtf(float t) {
#if defined(__CUDA_ARCH__)
    return hq_cbrtf_fast(t);
#else
    return hq_cbrtf_f64(t);
#endif
}

// ---------------------------------------------------------------- x / c for a constant c
// Correctly rounded quotient without the division unit (Markstein): q0 = RN(x*rc), the
// remainder x - q0*c is exact in one fma, q = RN(q0 + rem*rc).  rc = RN(1/c).  Verified
// exhaustively equal to x / c for every float with 2^-60 <= |x| <= 4 and the five constants the
// path divides by (both white points' X and Z, and 3*LABDELTA2) — tests/test_math_exhaustive.py;
// outside that range (zero, subnormal remainders, large values) the true division is used.
HQ_HD float hq_div_const(float x, float c, float rc) {
    const uint32_t ax = HQ_F2U(x) & 0x7fffffffu;
    if (ax - 0x21800000u <= 0x40800000u - 0x21800000u) {
        const float q0 = HQ_FMUL(x, rc);
        const float rem = HQ_FFMA(-q0, c, x);
        return HQ_FFMA(rem, rc, q0);
    }
    return HQ_FDIV(x, c);
}

// ---------------------------------------------------------------- (float)pow(b, 2.4f)
// Restates `(float)Math.pow(b, 2.4f)` (ScielabProcessor.java:282-284): the exponent
// is the float literal 2.4f widened to double, 2.4000000953674316.  b is a positive
// normal float (the sRGB decode calls it on (0.0901, 1.0]).
// ln b = k ln2 + 2 atanh(s), s = (m-1)/(m+1), carried as hi+lo; exp by n ln2 + f.
HQ_HD float hq_pow_2p4f(float b) {
    const double E = 0x1.3333340000000p+1;  // (double)2.4f
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    const double INV_LN2 = 1.44269504088896338700e+00;
    const double bd = (double)b;
    const uint64_t bits = HQ_D2U(bd);
    int k = (int)(bits >> 52) - 1023;
    double m = HQ_U2D((bits & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
    if (m > 1.4142135623730951) { m = HQ_DMUL(m, 0.5); k += 1; }
    const double num = HQ_DSUB(m, 1.0), den = HQ_DADD(m, 1.0);  // both exact
    const double s = HQ_DDIV(num, den);
    const double s_lo = HQ_DDIV(HQ_DFMA(-s, den, num), den);
    const double s2 = HQ_DMUL(s, s);
    double P = 1.0 / 23.0;
    P = HQ_DFMA(P, s2, 1.0 / 21.0);
    P = HQ_DFMA(P, s2, 1.0 / 19.0);
    P = HQ_DFMA(P, s2, 1.0 / 17.0);
    P = HQ_DFMA(P, s2, 1.0 / 15.0);
    P = HQ_DFMA(P, s2, 1.0 / 13.0);
    P = HQ_DFMA(P, s2, 1.0 / 11.0);
    P = HQ_DFMA(P, s2, 1.0 / 9.0);
    P = HQ_DFMA(P, s2, 1.0 / 7.0);
    P = HQ_DFMA(P, s2, 1.0 / 5.0);
    P = HQ_DFMA(P, s2, 1.0 / 3.0);
    const double kd = (double)k;
    const double a = HQ_DMUL(kd, LN2_HI), t2 = HQ_DMUL(2.0, s);
    const double hi = HQ_DADD(a, t2);
    const double err = HQ_DSUB(t2, HQ_DSUB(hi, a));
    const double tail = HQ_DADD(HQ_DMUL(2.0, s_lo), HQ_DMUL(t2, HQ_DMUL(s2, P)));
    const double lo = HQ_DADD(err, HQ_DADD(HQ_DMUL(kd, LN2_LO), tail));
    const double p_hi = HQ_DMUL(E, hi);
    const double p_lo = HQ_DADD(HQ_DFMA(E, hi, -p_hi), HQ_DMUL(E, lo));
    const double n = HQ_DRINT(HQ_DMUL(p_hi, INV_LN2));
    const double f_hi = HQ_DFMA(-n, LN2_HI, p_hi);
    const double f_lo = HQ_DFMA(-n, LN2_LO, p_lo);
    const double f = HQ_DADD(f_hi, f_lo);
    // exp(f) - 1 - f = f^2 * Q(f), Q = sum_{j=2..15} f^(j-2)/j!
    double Q = 1.0 / 1307674368000.0;
    Q = HQ_DFMA(Q, f, 1.0 / 87178291200.0);
    Q = HQ_DFMA(Q, f, 1.0 / 6227020800.0);
    Q = HQ_DFMA(Q, f, 1.0 / 479001600.0);
    Q = HQ_DFMA(Q, f, 1.0 / 39916800.0);
    Q = HQ_DFMA(Q, f, 1.0 / 3628800.0);
    Q = HQ_DFMA(Q, f, 1.0 / 362880.0);
    Q = HQ_DFMA(Q, f, 1.0 / 40320.0);
    Q = HQ_DFMA(Q, f, 1.0 / 5040.0);
    Q = HQ_DFMA(Q, f, 1.0 / 720.0);
    Q = HQ_DFMA(Q, f, 1.0 / 120.0);
    Q = HQ_DFMA(Q, f, 1.0 / 24.0);
    Q = HQ_DFMA(Q, f, 1.0 / 6.0);
    Q = HQ_DFMA(Q, f, 1.0 / 2.0);
    const double ff = HQ_DMUL(f, f);
    const double res = HQ_DADD(1.0, HQ_DADD(f_hi, HQ_DADD(f_lo, HQ_DMUL(ff, Q))));
    const double sc = HQ_U2D((uint64_t)(1023 + (int)n) << 52);
    return (float)HQ_DMUL(res, sc);
}

// ---------------------------------------------------------------- colour pipeline
// u8 -> [0,1]: the reference gets floats from Icy's convertToType(FLOAT, rescale)
// (HybridQuantization.java:95, third-party); the build defines it as (float)(c/255.0).
HQ_HD float hq_u8_to_unit(uint32_t c) { return (float)HQ_DDIV((double)c, 255.0); }

// ScielabProcessor.java:282-284
HQ_HD float hq_srgb_decode(float c) {
    if (c <= 0.04045f) return HQ_FDIV(c, 12.92f);
    return hq_pow_2p4f(HQ_FDIV(HQ_FADD(c, 0.055f), 1.055f));
}

// a*x + b*y + c*z evaluated left to right in fp32, every product and sum rounded
// (Java float arithmetic; no FMA).
HQ_HD float hq_dot3(float a, float x, float b, float y, float c, float z) {
    return HQ_FADD(HQ_FADD(HQ_FMUL(a, x), HQ_FMUL(b, y)), HQ_FMUL(c, z));
}

// ScielabProcessor.java:286-290 (linear RGB -> opponent)
HQ_HD hq_float3 hq_linrgb_to_opp(float R, float G, float B) {
    hq_float3 o;
    o.x = hq_dot3(0.26641335000823f, R, 0.60316740257478f, G, 0.0011333302293f, B);
    o.y = hq_dot3(-0.12197400229389f, R, 0.05598088396616f, G, 0.01326365114329f, B);
    o.z = hq_dot3(-0.08033445917708f, R, -0.33146741170125f, G, 0.44913244757774f, B);
    return o;
}

// ScielabProcessor.java:301 (one channel of XYZ -> Lab)
HQ_HD float hq_lab_f(float t) {
    if (t > HQ_LABDELTA3) return hq_cbrtf(t);
    return HQ_FADD(hq_div_const(t, HQ_3LABDELTA2, HQ_RCP_3LABDELTA2), HQ_4_OVER_29);
}
HQ_HD float hq_lab_f_f64(float t) {  // same with the fp64 cube root only (the device's re-check path)
    if (t > HQ_LABDELTA3) return hq_cbrtf_f64(t);
    return HQ_FADD(hq_div_const(t, HQ_3LABDELTA2, HQ_RCP_3LABDELTA2), HQ_4_OVER_29);
}

// white point with the reciprocals the constant division needs
struct hq_white {
    float x, y, z, rx, rz;
};
HQ_HD hq_white hq_make_white(int wp) {
    const hq_float3 w = hq_whitepoint(wp);
    hq_white o;
    o.x = w.x; o.y = w.y; o.z = w.z;
    o.rx = HQ_FDIV(1.0f, w.x); o.rz = HQ_FDIV(1.0f, w.z);
    return o;
}

// ScielabProcessor.java:295-310 (opponent -> XYZ -> Lab).  Y / illuminant[1] with illuminant[1] == 1.0f
// is Y itself for both white points (:20-21).
HQ_HD hq_float3 hq_opp_to_lab(hq_float3 o, hq_white white) {
    const float X = hq_dot3(0.97959616044562807864f, o.x, -1.5347157012664408981f, o.y,
                            0.44459764330437399288f, o.z);
    const float Y = hq_dot3(1.188977906742323787f, o.x, 0.7643549575179937615f, o.y,
                            0.13512574791125839373f, o.z);
    const float Z = hq_dot3(1.2318333139247290457f, o.x, 1.1631592597636512884f, o.y,
                            2.0784075888008567862f, o.z);
    const float fx = hq_lab_f(hq_div_const(X, white.x, white.rx));
    const float fy = hq_lab_f(white.y == 1.0f ? Y : HQ_FDIV(Y, white.y));
    const float fz = hq_lab_f(hq_div_const(Z, white.z, white.rz));
    hq_float3 lab;
    lab.x = HQ_FSUB(HQ_FMUL(116.0f, fy), 16.0f);
    lab.y = HQ_FMUL(500.0f, HQ_FSUB(fx, fy));
    lab.z = HQ_FMUL(200.0f, HQ_FSUB(fy, fz));
    return lab;
}

// linear RGB -> Lab (sRGBtoLab = OpptoLab(sRGBtoOpp(.)), ScielabProcessor.java:432)
HQ_HD hq_float3 hq_linrgb_to_lab(float R, float G, float B, hq_white white) {
    return hq_opp_to_lab(hq_linrgb_to_opp(R, G, B), white);
}
// the same through the fp64 cube root only: what rgb_to_lab_kernel falls back to for the pixels its packed
// fp32 fast path cannot decide
HQ_HD hq_float3 hq_linrgb_to_lab_f64(float R, float G, float B, hq_white white) {
    const hq_float3 o = hq_linrgb_to_opp(R, G, B);
    const float X = hq_dot3(0.97959616044562807864f, o.x, -1.5347157012664408981f, o.y, 0.44459764330437399288f, o.z);
    const float Y = hq_dot3(1.188977906742323787f, o.x, 0.7643549575179937615f, o.y, 0.13512574791125839373f, o.z);
    const float Z = hq_dot3(1.2318333139247290457f, o.x, 1.1631592597636512884f, o.y, 2.0784075888008567862f, o.z);
    const float fx = hq_lab_f_f64(hq_div_const(X, white.x, white.rx));
    const float fy = hq_lab_f_f64(white.y == 1.0f ? Y : HQ_FDIV(Y, white.y));
    const float fz = hq_lab_f_f64(hq_div_const(Z, white.z, white.rz));
    hq_float3 lab;
    lab.x = HQ_FSUB(HQ_FMUL(116.0f, fy), 16.0f);
    lab.y = HQ_FMUL(500.0f, HQ_FSUB(fx, fy));
    lab.z = HQ_FMUL(200.0f, HQ_FSUB(fy, fz));
    return lab;
}

// sRGB floats in [0,1] (a palette colour) -> Lab
HQ_HD hq_float3 hq_srgb_to_lab(float r, float g, float b, hq_white white) {
    return hq_linrgb_to_lab(hq_srgb_decode(r), hq_srgb_decode(g), hq_srgb_decode(b), white);
}

// Squared feature distance.  The reference compares OpenCL distance(float4,float4)
// (OptimizedConvolution.cl:180,185: implementation-defined precision, 4th lane 0-0).
// The build pins it: differences rounded to fp32, then dx*dx, fma(dy,dy,.), fma(dz,dz,.).
// Squared distances are compared (strict <, lowest index wins, :186) and the error
// of the winner is sqrt of that value, correctly rounded (CIE76, :209).
HQ_HD float hq_dist2(float x0, float x1, float x2, float p0, float p1, float p2) {
    const float d0 = HQ_FSUB(x0, p0), d1 = HQ_FSUB(x1, p1), d2 = HQ_FSUB(x2, p2);
    return HQ_FFMA(d2, d2, HQ_FFMA(d1, d1, HQ_FMUL(d0, d0)));
}

// ---------------------------------------------------------------- S-CIELAB stage (next row 1)
// The spatial-filter stage only exists as OpenCL kernels in the reference; their device-defined
// builtins are pinned as the oracle pins them: dot() left to right without contraction, fma()
// exact, pow/cbrt correctly rounded, distance() = sqrt of the fma chain.
//   RGB2XYZm cl:77, XYZ2Oppm cl:110, Opp2XYZm cl:118, RGB2Oppm cl:171
HQ_HD float hq_cl_dot3(float m0, float m1, float m2, float x, float y, float z) {
    return HQ_FADD(HQ_FADD(HQ_FMUL(x, m0), HQ_FMUL(y, m1)), HQ_FMUL(z, m2));
}
// linear RGB -> XYZ -> opponent (RGB2XYZ cl:79-90 + XYZ2Opp cl:111-116): the ORIGINAL image's route
HQ_HD hq_float3 hq_cl_linrgb_to_opp_via_xyz(float R, float G, float B) {
    const float X = hq_cl_dot3(0.4124564f, 0.3575761f, 0.1804375f, R, G, B);
    const float Y = hq_cl_dot3(0.2126729f, 0.7151522f, 0.0721750f, R, G, B);
    const float Z = hq_cl_dot3(0.0193339f, 0.1191920f, 0.9503041f, R, G, B);
    hq_float3 o;
    o.x = hq_cl_dot3(0.2787336f, 0.7218031f, -0.1065520f, X, Y, Z);
    o.y = hq_cl_dot3(-0.4487736f, 0.2898056f, -0.0771569f, X, Y, Z);
    o.z = hq_cl_dot3(0.0859513f, -0.5899859f, 0.5011089f, X, Y, Z);
    return o;
}
// linear RGB -> opponent directly (quantizeAndConvertToOpp cl:194-198): the QUANTISED image's route
HQ_HD hq_float3 hq_cl_linrgb_to_opp(float R, float G, float B) {
    hq_float3 o;
    o.x = hq_cl_dot3(0.266413f, 0.603167f, 0.00113333f, R, G, B);
    o.y = hq_cl_dot3(-0.124957f, 0.0375879f, -0.133381f, R, G, B);
    o.z = hq_cl_dot3(-0.0803345f, -0.331467f, 0.449132f, R, G, B);
    return o;
}
// Opp2LAB cl:124-145: LABDELTA3 = 216f/24389f (= 0x1.22354ep-7f, the same float as Java's),
// kappa = 24389f/27f, f = t > LABDELTA3 ? cbrt(t) : fma(kappa, t, 16)/116
HQ_HD float hq_cl_lab_f(float t) {
    if (t > HQ_LABDELTA3) return hq_cbrtf(t);
    return HQ_FDIV(HQ_FFMA(0x1.c3a5eep+9f, t, 16.0f), 116.0f);
}
HQ_HD hq_float3 hq_cl_opp_to_lab(float o0, float o1, float o2, hq_float3 ill) {
    const float X = hq_cl_dot3(0.624045f, -1.87044f, -0.155304f, o0, o1, o2);
    const float Y = hq_cl_dot3(1.36606f, 0.931563f, 0.433903f, o0, o1, o2);
    const float Z = hq_cl_dot3(1.5013f, 1.41761f, 2.53307f, o0, o1, o2);
    const float fx = hq_cl_lab_f(HQ_FDIV(X, ill.x));
    const float fy = hq_cl_lab_f(HQ_FDIV(Y, ill.y));
    const float fz = hq_cl_lab_f(HQ_FDIV(Z, ill.z));
    hq_float3 lab;
    lab.x = HQ_FSUB(HQ_FMUL(116.0f, fy), 16.0f);
    lab.y = HQ_FMUL(500.0f, HQ_FSUB(fx, fy));
    lab.z = HQ_FMUL(200.0f, HQ_FSUB(fy, fz));
    return lab;
}
// the same with the illuminant known to be one of the plugin's two white points (ScielabProcessor.java:20-21; Y = 1):
// X / Xn and Z / Zn by the exhaustively verified constant division (hq_div_const == x / c for every float), Y / 1 = Y
HQ_HD hq_float3 hq_cl_opp_to_lab_white(float o0, float o1, float o2, hq_white w) {
    const float X = hq_cl_dot3(0.624045f, -1.87044f, -0.155304f, o0, o1, o2);
    const float Y = hq_cl_dot3(1.36606f, 0.931563f, 0.433903f, o0, o1, o2);
    const float Z = hq_cl_dot3(1.5013f, 1.41761f, 2.53307f, o0, o1, o2);
    const float fx = hq_cl_lab_f(hq_div_const(X, w.x, w.rx));
    const float fy = hq_cl_lab_f(Y);
    const float fz = hq_cl_lab_f(hq_div_const(Z, w.z, w.rz));
    hq_float3 lab;
    lab.x = HQ_FSUB(HQ_FMUL(116.0f, fy), 16.0f);
    lab.y = HQ_FMUL(500.0f, HQ_FSUB(fx, fy));
    lab.z = HQ_FMUL(200.0f, HQ_FSUB(fy, fz));
    return lab;
}
// CIEDE kernel, CIE94 branch (cl:217-226; scope row f4).  p1 = the original's Lab, p2 = the other image's.  OpenCL's device-
// defined pieces pinned as the oracle pins them: fma exact, sqrt and the divisions correctly rounded, the weights
// 1 + 0.045*C1 and 1 + 0.015*C1 evaluated in double (the literals are doubles in OpenCL C) and narrowed on assignment.
// deltaH takes the square root of da^2 + db^2 - dC^2, which rounding makes slightly negative for collinear chroma vectors:
// the result is then NaN, exactly as the reference kernel's (about one generic pixel pair in 4,000).
HQ_HD float hq_cl_delta_e94(float L1, float a1, float b1, float L2, float a2, float b2) {
    const float dL = HQ_FSUB(L1, L2);
    const float c1 = HQ_FSQRT(HQ_FFMA(a1, a1, HQ_FMUL(b1, b1)));
    const float dC = HQ_FSUB(c1, HQ_FSQRT(HQ_FFMA(a2, a2, HQ_FMUL(b2, b2))));
    const float da = HQ_FSUB(a1, a2), db = HQ_FSUB(b1, b2);
    const float dH = HQ_FSQRT(HQ_FSUB(HQ_FFMA(da, da, HQ_FMUL(db, db)), HQ_FMUL(dC, dC)));
    const float sc = (float)HQ_DADD(1.0, HQ_DMUL(0.045, (double)c1));
    const float sh = (float)HQ_DADD(1.0, HQ_DMUL(0.015, (double)c1));
    const float qc = HQ_FDIV(dC, sc), qh = HQ_FDIV(dH, sh);
    return HQ_FSQRT(HQ_FFMA(dL, dL, HQ_FFMA(qc, qc, HQ_FMUL(qh, qh))));
}
// reflect padding of the separable filters (cl:20-27)
HQ_HD int hq_reflect(int off, int n) {
    if (off < 0) off = -off - 1;
    else if (off >= n) off = (n << 1) - off - 1;
    return off;
}

// test hook: the restated routines by code (0 cube root, 1 pow 2.4f, 2 sRGB decode,
// 3/4 x / Xn, x / Zn of D65, 5 x / (3*LABDELTA2), 6/7 x / Xn, x / Zn of D50)
HQ_HD float hq_math_probe(int which, float v) {
    if (which == 0) return hq_cbrtf(v);
    if (which == 1) return hq_pow_2p4f(v);
    if (which == 2) return hq_srgb_decode(v);
    if (which == 5) return hq_div_const(v, HQ_3LABDELTA2, HQ_RCP_3LABDELTA2);
    const hq_white w = hq_make_white(which >= 6 ? HQ_WHITE_D50 : HQ_WHITE_D65);
    return (which == 3 || which == 6) ? hq_div_const(v, w.x, w.rx) : hq_div_const(v, w.z, w.rz);
}

// float -> 2^-24 fixed point, round to nearest even (exact for |v| >= 1)
HQ_HD long long hq_to_fx(float v) { return HQ_F2LL_RN(HQ_FMUL(v, HQ_FX_SCALE)); }
