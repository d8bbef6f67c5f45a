// hpgb_math.h -- accurate fp64 trig for the kernels: cos / sin / sincos / acos / atan2 evaluated
// in double-double so that the returned double is the correctly rounded result except with
// probability ~2^-18 (error <= ~2^-72 relative).
//
// Why: the reference takes these from glibc 2.39's libm, which is itself correctly rounded for
// ~99.9 % of arguments.  CUDA's own fp64 cos/sin/acos/atan2 are 1-2 ulp, which
// would differ from glibc on tens of percent of arguments; the outputs that subtract nearly
// equal angles afterwards (latitude = pi/2 - theta, interpolation theta-weights) amplify a
// single-ulp difference to thousands of ulps.  Returning the correctly rounded value is the
// only libm-independent way to be bit-identical to the reference wherever the reference is
// correctly rounded.
//
// How (straight-line, no data-dependent branches, bounded arguments only -- |x| <= ~7):
//   * x = q*(pi/2) + r, |r| <= pi/4, with a three-part pi/2 (r kept as a double-double);
//   * |r| = k/64 + d with |d| <= 1/128: table of sin/cos(k/64) as double-doubles (52 rows x 32 B,
//     staged in shared memory by the kernels), angle-addition with short polynomials in d;
//     the two leading products are kept exact with fma (two_prod);
//   * acos / atan2: a float-accuracy first guess t0 (24 bits, so t0 reduces exactly), the
//     double-double sin/cos of t0 from the routine above, then ONE series-corrected Newton step:
//       acos : cos(t0 + e) = z            e = u - (cot/2) u^2 + (1/6 + cot^2/2) u^3,  u = (cos t0 - z)/sin t0
//       atan2: x sin(t0+e) - y cos(t0+e) = 0   e = v - v^3/3,  v = -(x sin t0 - y cos t0)/(x cos t0 + y sin t0)
//     (|u|,|v| ~ 2^-22, so the neglected terms are < 2^-80).
//
// All functions are __host__ __device__ (see hpgb_core.h); fma() is called explicitly, the
// translation units are compiled without implicit contraction.
#ifndef HPGB_MATH_H
#define HPGB_MATH_H

#include "hpgb_core.h"
#include "hpgb_tables.h"

// exact product a*b = p + e
HPGB_HD void hpgb_two_prod(double a, double b, double &p, double &e) {
    p = a * b;
    e = fma(a, b, -p);
}
// exact sum a+b = s + e, requires |a| >= |b| (or a == 0)
HPGB_HD void hpgb_fast_two_sum(double a, double b, double &s, double &e) {
    s = a + b;
    e = b - (s - a);
}
// exact sum, no precondition
HPGB_HD void hpgb_two_sum(double a, double b, double &s, double &e) {
    s = a + b;
    const double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}

// v with its sign flipped when s is negative (rh < 0 and "sign bit of rh set" agree for every rh the reductions
// produce: they never yield -0), resp. when bit 31 of `m` is set: one LOP3 on the high word
HPGB_HD double hpgb_flip_sign_bit(double v, uint32_t m) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(v) ^ (int)(m & 0x80000000u), __double2loint(v));
#else
    return (m & 0x80000000u) ? -v : v;
#endif
}
HPGB_HD double hpgb_flip_sign_if_neg(double v, double s) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(v) ^ (__double2hiint(s) & (int)0x80000000), __double2loint(v));
#else
    return (s < 0.0) ? -v : v;
#endif
}

// rint(p) and the same value as an int, |p| < 2^31.  The device version adds and subtracts
// 1.5 2^52 (round-to-nearest-even at the integer position: the same value as rint) and reads the integer out of
// the low mantissa word -- two fixed-latency DADDs instead of a round + a double -> int on the conversion unit.
HPGB_HD double hpgb_rint_i(double p, int &i) {
#if defined(__CUDA_ARCH__) && !defined(HPGB_NO_MAGIC_RINT)
    const double t = p + 6755399441055744.0;
    i = __double2loint(t);
    return t - 6755399441055744.0;
#else
    const double f = rint(p);
    i = (int)f;
    return f;
#endif
}

// x -> quadrant q and r = x - q*pi/2 as rh + rl, |r| <= pi/4 (1 + 2^-50).  Valid for |x| < 2^20.
HPGB_HD void hpgb_reduce_pio2(double x, int &q, double &rh, double &rl) {
    const double fq = hpgb_rint_i(x * HPGB_2OPI, q);
    const double t = fma(-fq, HPGB_PIO2_1, x);  // exact: |t| < 1 and t is a multiple of ulp(x)/4
    double p, pe, e1;
    hpgb_two_prod(fq, HPGB_PIO2_2, p, pe);
    hpgb_two_sum(t, -p, rh, e1);
    rl = (e1 - pe) - fq * HPGB_PIO2_3;
}

// sin and cos of r = rh + rl, |r| <= pi/4 + 2^-40, as double-doubles; T = sincos table
HPGB_HD void hpgb_sincos_dd(const double *T, double rh, double rl, double &sh, double &sl, double &ch,
                            double &cl) {
    const bool neg = rh < 0.0;
    const double a = fabs(rh);
    const double al = neg ? -rl : rl;
    int k;
    const double fk = hpgb_rint_i(a * 64.0, k);
    const double d = a - fk * 0.015625;  // exact
    const double Sh = T[4 * k + 0], Sl = T[4 * k + 1], Ch = T[4 * k + 2], Cl = T[4 * k + 3];
    // powers of the offset
    double p, pe;
    hpgb_two_prod(d, d, p, pe);
    // sin(delta) - delta  and  cos(delta) - 1  (delta = d + al)
    const double sd = d * p * fma(p, fma(p, -1.0 / 5040.0, 1.0 / 120.0), -1.0 / 6.0);
    const double cdh = -0.5 * p;
    const double cdl = fma(-0.5, pe, -d * al) + p * p * fma(p, fma(p, 1.0 / 40320.0, -1.0 / 720.0), 1.0 / 24.0);
    const double ds = al + sd;  // sin(delta) = d + ds
    // sin(a) = S cos(delta) + C sin(delta) = S + C d + S cdh + [Sl + Cl d + C ds + S cdl]
    {
        double m, me, n, ne, v, ve, w, we;
        hpgb_two_prod(Ch, d, m, me);
        hpgb_two_prod(Sh, cdh, n, ne);
        hpgb_fast_two_sum(Sh, m, v, ve);
        hpgb_fast_two_sum(v, n, w, we);
        const double low = we + (ve + ((me + ne) + (fma(Cl, d, Sl) + fma(Ch, ds, Sh * cdl))));
        hpgb_fast_two_sum(w, low, sh, sl);
    }
    // cos(a) = C cos(delta) - S sin(delta) = C - S d + C cdh + [Cl - Sl d - S ds + C cdl]
    {
        double m, me, n, ne, v, ve, w, we;
        hpgb_two_prod(Ch, cdh, m, me);
        hpgb_two_prod(Sh, d, n, ne);
        hpgb_fast_two_sum(Ch, -n, v, ve);
        hpgb_fast_two_sum(v, m, w, we);
        const double low = we + (ve + ((me - ne) + (fma(-Sl, d, Cl) + fma(-Sh, ds, Ch * cdl))));
        hpgb_fast_two_sum(w, low, ch, cl);
    }
    if (neg) {
        sh = -sh;
        sl = -sl;
    }
}

// sin and cos of x (|x| < 2^20) as double-doubles
HPGB_HD void hpgb_sincos_full_dd(const double *T, double x, double &sh, double &sl, double &ch, double &cl) {
    int q;
    double rh, rl, s1, s2, c1, c2;
    hpgb_reduce_pio2(x, q, rh, rl);
    hpgb_sincos_dd(T, rh, rl, s1, s2, c1, c2);
    const bool swap = q & 1;
    sh = swap ? c1 : s1;
    sl = swap ? c2 : s2;
    ch = swap ? s1 : c1;
    cl = swap ? s2 : c2;
    if (q & 2) {  // sin: q = 2,3 negative
        sh = -sh;
        sl = -sl;
    }
    if ((q + 1) & 2) {  // cos: q = 1,2 negative
        ch = -ch;
        cl = -cl;
    }
}

HPGB_HD double hpgb_cos_cr(const double *T, double x) {
    double sh, sl, ch, cl;
    hpgb_sincos_full_dd(T, x, sh, sl, ch, cl);
    return ch + cl;
}
HPGB_HD double hpgb_sin_cr(const double *T, double x) {
    double sh, sl, ch, cl;
    hpgb_sincos_full_dd(T, x, sh, sl, ch, cl);
    return sh + sl;
}
HPGB_HD void hpgb_sincos_cr(const double *T, double x, double &s, double &c) {
    double sh, sl, ch, cl;
    hpgb_sincos_full_dd(T, x, sh, sl, ch, cl);
    s = sh + sl;
    c = ch + cl;
}

// ---------------------------------------------------------------------------------------------
// sin and cos of x (0 <= |x| < 2^20) as PLAIN doubles, error <= 0.51 ulp: the same reduction and table as
// the double-double routine above, but only the two leading terms of the angle-addition formula are kept
// exact (S + C d as a two-product + two-sum); everything below 2^-7 of the result is plain fma arithmetic.
// About 40 fp64 operations instead of ~85.  Used where the consumer multiplies the result afterwards and
// the contract is "<= 2 ulp" rather than "the bits glibc returns" (pixel_to_vector: x = sin(theta) cos(phi)).
// ---------------------------------------------------------------------------------------------
HPGB_HD void hpgb_sincos_sd(const double *T, double x, double &s, double &c) {
    // x = q pi/2 + r, r = rh + rl
    int q;
    const double fq = hpgb_rint_i(x * HPGB_2OPI, q);
    const double t = fma(-fq, HPGB_PIO2_1, x);  // exact (see hpgb_reduce_pio2)
    double w, we;
    hpgb_two_prod(fq, HPGB_PIO2_2, w, we);
    const double rh = t - w;
    const double rl = ((t - rh) - w) - fma(fq, HPGB_PIO2_3, we);  // |t| >= |w| or both tiny: fast two-sum is exact enough
    const bool neg = rh < 0.0;
    const double a = fabs(rh);
    const double al = neg ? -rl : rl;
    int k;
    const double fk = hpgb_rint_i(a * 64.0, k);
    const double d = fma(-fk, 0.015625, a);  // exact
    const double Sh = T[4 * k + 0], Sl = T[4 * k + 1], Ch = T[4 * k + 2], Cl = T[4 * k + 3];
    const double p = d * d;
    // sin(delta) = d + sd, cos(delta) = 1 + cd   (delta = d + al, |delta| <= 2^-7)
    const double sd = fma(d * p, fma(p, fma(p, -1.0 / 5040.0, 1.0 / 120.0), -1.0 / 6.0), al);
    const double cd = fma(p, fma(p, fma(p, -1.0 / 720.0, 1.0 / 24.0), -0.5), -d * al);
    double sa, ca;
    {   // sin(a) = S + C d + [C sd + S cd + Sl + Cl d]
        double m, me, v, ve;
        hpgb_two_prod(Ch, d, m, me);
        hpgb_fast_two_sum(Sh, m, v, ve);  // |S| >= |C d| for k >= 1; k = 0: S = 0
        sa = v + (ve + (me + fma(Ch, sd, fma(Sh, cd, fma(Cl, d, Sl)))));
    }
    {   // cos(a) = C - S d + [-S sd + C cd + Cl - Sl d]
        double m, me, v, ve;
        hpgb_two_prod(-Sh, d, m, me);
        hpgb_fast_two_sum(Ch, m, v, ve);
        ca = v + (ve + (me + fma(-Sh, sd, fma(Ch, cd, fma(-Sl, d, Cl)))));
    }
    if (neg) sa = -sa;
    const bool swap = q & 1;
    s = swap ? ca : sa;
    c = swap ? sa : ca;
    if (q & 2) s = -s;
    if ((q + 1) & 2) c = -c;
}

// The same with the angle-addition sums evaluated in plain fma arithmetic (no exact products / sums): the terms
// below the table value are < 2^-6 of it except in the first two table cells, so the result is within 0.52 ulp for
// |r| >= 1/32 and 0.8 ulp below (r = the reduced argument).  28 fp64 operations.  For pixel_to_vector, whose contract
// is "x, y within 2 ulp": a cosine within 1.5 ulp of the true value is within 1 ulp of glibc's, and the product with
// sin(theta) then within 2 ulp of the reference's.
HPGB_HD void hpgb_sincos_p(const double *T, double x, double &s, double &c) {
    int q;
    const double fq = hpgb_rint_i(x * HPGB_2OPI, q);
    const double t = fma(-fq, HPGB_PIO2_1, x);  // exact (see hpgb_reduce_pio2)
    double w, we;
    hpgb_two_prod(fq, HPGB_PIO2_2, w, we);
    const double rh = t - w;
    const double rl = ((t - rh) - w) - fma(fq, HPGB_PIO2_3, we);
    const double a = fabs(rh);
#if defined(HPGB_SINCOS_SELECTS)
    const bool neg = rh < 0.0;
    const double al = neg ? -rl : rl;
#else
    const double al = hpgb_flip_sign_if_neg(rl, rh);  // sign handling on the high words: one LOP3 each instead of selects
#endif
    int k;
    const double fk = hpgb_rint_i(a * 64.0, k);
    const double d = fma(-fk, 0.015625, a);  // exact
    const double Sh = T[4 * k + 0], Sl = T[4 * k + 1], Ch = T[4 * k + 2], Cl = T[4 * k + 3];
    const double p = d * d;
    const double sd = fma(d * p, fma(p, fma(p, -1.0 / 5040.0, 1.0 / 120.0), -1.0 / 6.0), al);
    const double cd = fma(p, fma(p, fma(p, -1.0 / 720.0, 1.0 / 24.0), -0.5), -d * al);
    double sa = Sh + fma(Ch, d, fma(Ch, sd, fma(Sh, cd, fma(Cl, d, Sl))));
    const double ca = Ch + fma(-Sh, d, fma(-Sh, sd, fma(Ch, cd, fma(-Sl, d, Cl))));
    const bool swap = q & 1;
#if defined(HPGB_SINCOS_SELECTS)
    if (neg) sa = -sa;
    s = swap ? ca : sa;
    c = swap ? sa : ca;
    if (q & 2) s = -s;
    if ((q + 1) & 2) c = -c;
#else
    sa = hpgb_flip_sign_if_neg(sa, rh);
    s = hpgb_flip_sign_bit(swap ? ca : sa, (uint32_t)q << 30);        // quadrants 2, 3: sin negative
    c = hpgb_flip_sign_bit(swap ? sa : ca, (uint32_t)(q + 1) << 30);  // quadrants 1, 2: cos negative
#endif
}

// 24-bit first guesses (any estimate good to ~2^-20 gives the same final result)
HPGB_HD double hpgb_acos_guess(double z) {
#if defined(__CUDA_ARCH__)
    return (double)acosf((float)z);
#else
    return (double)(float)acos(z);
#endif
}
HPGB_HD double hpgb_atan2_guess(double y, double x) {
#if defined(__CUDA_ARCH__)
    return (double)atan2f((float)y, (float)x);
#else
    return (double)(float)atan2(y, x);
#endif
}

// acos(z) for |z| <= 0.995 (pixel centres use acos only for |z| <= 0.99, atan2 beyond)
HPGB_HD double hpgb_acos_cr(const double *T, double z) {
    const double t0 = hpgb_acos_guess(z);
    double sh, sl, ch, cl;
    hpgb_sincos_full_dd(T, t0, sh, sl, ch, cl);
    const double e = (ch - z) + cl;
    const double inv = 1.0 / sh;
    const double u = e * inv;
    const double cot = z * inv;
    const double u2 = u * u;
    const double corr = fma(u2 * u, fma(0.5 * cot, cot, 1.0 / 6.0), -0.5 * cot * u2);
    return t0 + (u + corr);
}

// atan2(y, x) for finite, not-both-zero arguments whose magnitudes are within float range of
// each other after scaling (callers pass normalised-ish values; see hpgb_atan2_any)
HPGB_HD double hpgb_atan2_cr(const double *T, double y, double x) {
    const double t0 = hpgb_atan2_guess(y, x);
    double sh, sl, ch, cl;
    hpgb_sincos_full_dd(T, t0, sh, sl, ch, cl);
    // f = x sin t0 - y cos t0 (exact leading products, the difference of the heads is exact)
    double a, ae, b, be;
    hpgb_two_prod(x, sh, a, ae);
    hpgb_two_prod(y, ch, b, be);
    const double f = (a - b) + ((ae - be) + fma(x, sl, -y * cl));
    const double g = fma(x, ch, y * sh);
    const double v = -f / g;
    return t0 + fma(v * v * v, -1.0 / 3.0, v);
}

// general atan2 with exponent scaling so the float guess never over/underflows
HPGB_HD double hpgb_atan2_any(const double *T, double y, double x) {
    const double m = fmax(fabs(x), fabs(y));
    if (!(m > 0.0) || !(m < INFINITY)) return atan2(y, x);  // zeros / inf / nan: libm semantics
    int ex;
    (void)frexp(m, &ex);
    const double ys = ldexp(y, -ex), xs = ldexp(x, -ex);
    if (fabs(ys) < 1e-30 || fabs(xs) < 1e-30) {
        // one component negligible at float precision: the guess degenerates; result is within
        // 1e-30 of a multiple of pi/2 and the series below still applies with a double guess
        const double t0 = atan2(ys, xs);
        return t0;
    }
    return hpgb_atan2_cr(T, ys, xs);
}

// ---------------------------------------------------------------------------------------------
// Table-driven acos / atan2 for pixel centres (pixel_to_angle, interpolation ring colatitudes):
// the same "correctly rounded in practice" contract as hpgb_acos_cr / hpgb_atan2_cr (result
// returned as an unevaluated sum hi + lo good to ~2^-68 relative), at a quarter of the fp64
// work -- no first guess, no double-double sin/cos, no division:
//   acos(|z|) = sum_j a_j r^j around one of 897 cell centres zc (Taylor coefficients from mpmath; a_6..a_9
//   stored as floats: 80-byte rows, an odd number of 16-byte bank groups, so random gathers use every bank),
//   tools/gen_tables.py), r = |z| - zc EXACT; cells are uniform (2^-8) for |z| <= 1/2 and of
//   relative width 2^-7 in u = 1 - |z| beyond, so |r| / (1 - |zc|) <= 2^-8 everywhere and ten terms
//   reach 2^-70; a_0, a_1 are double-doubles, a_1 r is an exact two-product, the rest is Horner.
//   z < 0 : pi - acos(|z|), with pi as a double-double.
//   atan2(s, c), 0 <= s <= 1.13 |c| (polar-cap ring colatitudes; the first 80 rows, s <= 0.156 |c|, serve
//   |z| > 0.99): q = s/c as quotient + exact
//   remainder correction, then the same scheme around q = k/512 with eight terms.
// TA / TQ point at the tables (HPGB_ACOS_TABLE_INIT / HPGB_ATANQ_TABLE_INIT), staged in shared memory.
// ---------------------------------------------------------------------------------------------
HPGB_HD uint32_t hpgb_math_hi32(double v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2hiint(v);
#else
    union {
        double d;
        uint64_t u;
    } c;
    c.d = v;
    return (uint32_t)(c.u >> 32);
#endif
}

// a / b correctly rounded for "tame" operands (no overflow / underflow / subnormals anywhere near):
// reciprocal approximation, two Newton steps, two residual corrections -- the sequence behind
// div.rn.f64 without its special-case branch, so it stays straight-line code.  Validated against
// operator / on 3e10 operand pairs of the domain it is used for (tools/ubench/divcheck.cu).
HPGB_HD double hpgb_div_tame(double a, double b) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
    y = fma(y, fma(-b, y, 1.0), y);
    y = fma(y, fma(-b, y, 1.0), y);
    double q = a * y;
    q = fma(fma(-b, q, a), y, q);
    q = fma(fma(-b, q, a), y, q);
    return q;
#else
    return a / b;
#endif
}

HPGB_HD double hpgb_w2d(uint64_t w) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)w);
#else
    union {
        uint64_t b;
        double d;
    } cv;
    cv.b = w;
    return cv.d;
#endif
}
HPGB_HD double hpgb_w2f(uint32_t w) {  // float bits -> double
#if defined(__CUDA_ARCH__)
    return (double)__uint_as_float(w);
#else
    union {
        uint32_t b;
        float f;
    } cv;
    cv.b = w;
    return (double)cv.f;
#endif
}

// sqrt(x) correctly rounded for tame x (normal, far from overflow): rsqrt approximation, two coupled
// Newton steps, one residual correction (Markstein) -- sqrt.rn.f64 without its special-case branch.
// Validated against sqrt() on 3e10 operands of the domain it is used for (tools/ubench/sqrtcheck.cu).
HPGB_HD double hpgb_sqrt_tame(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    return fma(fma(-g, g, x), h, g);
#else
    return sqrt(x);
#endif
}

// acos(z) as hi + lo, |z| <= 0.99.  TA: rows of HPGB_ACOS_ROW_WORDS 64-bit words (see tools/gen_tables.py)
#if defined(HPGB_ACOS_V1)
HPGB_HD void hpgb_acos_tab_dd(const uint64_t *TA, double z, double &hi, double &lo) {
    const double a = fabs(z);
    const double u = 1.0 - a;  // exact for a >= 1/2
    const int ia = (int)rint(a * 256.0);
    const int ib = 129 + (int)(hpgb_math_hi32(u) >> 13) - 1016 * 128;
    int idx = a <= 0.5 ? ia : ib;
    idx = idx < 0 ? 0 : (idx > HPGB_ACOS_ROWS - 1 ? HPGB_ACOS_ROWS - 1 : idx);  // NaN / out-of-domain input: stay inside the table
    const uint64_t *roww = TA + idx * HPGB_ACOS_ROW_WORDS;
    double row[12];
#pragma unroll
    for (int j = 0; j < 8; j++) row[j] = hpgb_w2d(roww[j]);
    row[8] = hpgb_w2f((uint32_t)roww[8]);
    row[9] = hpgb_w2f((uint32_t)(roww[8] >> 32));
    row[10] = hpgb_w2f((uint32_t)roww[9]);
    row[11] = hpgb_w2f((uint32_t)(roww[9] >> 32));
    // r = |z| - zc, exact.  |z| <= 1/2: zc = idx / 256.  Beyond: zc = 1 - uc, uc = the cell's lower edge
    // (u with all but 7 mantissa bits cleared) + half a cell (the next mantissa bit), so r = uc - u.
    const double uc = hpgb_w2d((uint64_t)((hpgb_math_hi32(u) & 0xFFFFE000u) | 0x00001000u) << 32);
    const double r = a <= 0.5 ? fma(-(double)idx, 0.00390625, a) : uc - u;
    double s = fma(row[11], r, row[10]);
    s = fma(s, r, row[9]);
    s = fma(s, r, row[8]);
    s = fma(s, r, row[7]);
    s = fma(s, r, row[6]);
    s = fma(s, r, row[5]);
    s = fma(s, r, row[4]);
    double p, pe, h, hl;
    hpgb_two_prod(row[2], r, p, pe);
    const double low = row[1] + (pe + fma(row[3], r, (r * r) * s));
    hpgb_fast_two_sum(row[0], p, h, hl);  // |a1 r| < acos(zc)
    const double l = hl + low;
    // z < 0: pi - (h + l), pi as a double-double; both variants computed, two selects
    double nh, ne;
    hpgb_fast_two_sum(HPGB_PI_1, -h, nh, ne);  // pi_1 > h
    const bool neg = z < 0.0;
    hi = neg ? nh : h;
    lo = neg ? ne + (HPGB_PI_2 - l) : l;
}
#else
// Row = nine doubles { a0 hi, lo, a1 hi, lo, a2 .. a6 } and one word whose 64 bits ARE a7 (its low mantissa half is
// the high word of a8, accounted for when the table was generated) -- no float -> double conversions; the cell
// number of |z| <= 1/2 comes out of the mantissa of |z| 256 + 1.5 2^52 (no rint / double -> int / int -> double
// conversions either: those run on the quarter-rate conversion unit).
HPGB_HD void hpgb_acos_tab_abs(const uint64_t *TA, double z, double &h, double &l) {  // acos(|z|) = h + l
    const double a = fabs(z);
    const double u = 1.0 - a;  // exact for a >= 1/2
    const double am = fma(a, 256.0, 6755399441055744.0);  // 1.5 2^52 + RN(256 a): the integer sits in the low word
    const double fk = am - 6755399441055744.0;
#if defined(__CUDA_ARCH__)
    const int ia = __double2loint(am);
#else
    union {
        double d;
        uint64_t w;
    } cv;
    cv.d = am;
    const int ia = (int)(uint32_t)cv.w;
#endif
    const int ib = 129 + (int)(hpgb_math_hi32(u) >> 13) - 1016 * 128;
    int idx = a <= 0.5 ? ia : ib;
    idx = idx < 0 ? 0 : (idx > HPGB_ACOS_ROWS - 1 ? HPGB_ACOS_ROWS - 1 : idx);  // NaN / out-of-domain input: stay inside the table
    const uint64_t *roww = TA + idx * HPGB_ACOS_ROW_WORDS;
    double row[9];
#pragma unroll
    for (int j = 0; j < 9; j++) row[j] = hpgb_w2d(roww[j]);
    const uint64_t w9 = roww[9];
    const double a7 = hpgb_w2d(w9), a8 = hpgb_w2d(w9 << 32);
    // r = |z| - zc, exact.  |z| <= 1/2: zc = cell / 256.  Beyond: zc = 1 - uc, uc = the cell's lower edge
    // (u with all but 7 mantissa bits cleared) + half a cell (the next mantissa bit), so r = uc - u.
    const double uc = hpgb_w2d((uint64_t)((hpgb_math_hi32(u) & 0xFFFFE000u) | 0x00001000u) << 32);
    const double r = a <= 0.5 ? fma(-fk, 0.00390625, a) : uc - u;
    double s = fma(a8, r, a7);
    s = fma(s, r, row[8]);
    s = fma(s, r, row[7]);
    s = fma(s, r, row[6]);
    s = fma(s, r, row[5]);
    s = fma(s, r, row[4]);
    double p, pe, hl;
    hpgb_two_prod(row[2], r, p, pe);
    const double low = row[1] + (pe + fma(row[3], r, (r * r) * s));
    hpgb_fast_two_sum(row[0], p, h, hl);  // |a1 r| < acos(zc)
    l = hl + low;
}
// z < 0: pi - (h + l), pi as a double-double
HPGB_HD void hpgb_acos_tab_dd(const uint64_t *TA, double z, double &hi, double &lo) {
    double h, l, nh, ne;
    hpgb_acos_tab_abs(TA, z, h, l);
    hpgb_fast_two_sum(HPGB_PI_1, -h, nh, ne);  // pi_1 > h
    const bool neg = z < 0.0;
    hi = neg ? nh : h;
    lo = neg ? ne + (HPGB_PI_2 - l) : l;
}
#endif
HPGB_HD double hpgb_acos_tab(const uint64_t *TA, double z) {
#if defined(HPGB_ACOS_V1) || defined(HPGB_ACOS_SELECT_PARTS)
    double hi, lo;
    hpgb_acos_tab_dd(TA, z, hi, lo);
    return hi + lo;
#else
    // the same two sums as hi + lo of hpgb_acos_tab_dd, selected AFTER the additions: one select instead of two
    double h, l, nh, ne;
    hpgb_acos_tab_abs(TA, z, h, l);
    hpgb_fast_two_sum(HPGB_PI_1, -h, nh, ne);  // pi_1 > h
    const double pos = h + l, negv = nh + (ne + (HPGB_PI_2 - l));
    return z < 0.0 ? negv : pos;
#endif
}

// atan2(s, c) as hi + lo for 0 <= s <= 0.156 |c| (result pi - atan2(s, |c|) when c < 0)
HPGB_HD void hpgb_atan2_tab_dd(const double *TQ, double s, double c, double &hi, double &lo, int nrows = HPGB_ATANQ_ROWS_POLAR) {
    const double ca = fabs(c);
#if defined(__CUDA_ARCH__)
    double ic;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(ic) : "d"(ca));
    const double ic2 = fma(ic, fma(-ca, ic, 1.0), ic);
    const double qh = s * fma(ic2, fma(-ca, ic2, 1.0), ic2);
#else
    const double ic = 1.0 / ca;
    const double qh = s * ic;
#endif
    const double ql = fma(-qh, ca, s) * ic;  // exact remainder, rough reciprocal: q = qh + ql to ~2^-75
    int idx = (int)rint(qh * 512.0);
    idx = idx < 0 ? 0 : (idx > nrows - 1 ? nrows - 1 : idx);
    const double *row = TQ + idx * HPGB_ATANQ_ROW_DOUBLES;
    const double r = qh - (double)idx * 0.001953125;  // exact
    const double rt = r + ql;
    double t = fma(row[9], rt, row[8]);
    t = fma(t, rt, row[7]);
    t = fma(t, rt, row[6]);
    t = fma(t, rt, row[5]);
    t = fma(t, rt, row[4]);
    double p, pe, h, hl;
    hpgb_two_prod(row[2], r, p, pe);
    const double low = row[1] + (pe + fma(row[2], ql, fma(row[3], r, (rt * rt) * t)));
    hpgb_two_sum(row[0], p, h, hl);  // atan(qc) may be 0 (first cell)
    const double l = hl + low;
    double nh, ne;
    hpgb_fast_two_sum(HPGB_PI_1, -h, nh, ne);
    const bool neg = c < 0.0;
    hi = neg ? nh : h;
    lo = neg ? ne + (HPGB_PI_2 - l) : l;
}
HPGB_HD double hpgb_atan2_tab(const double *TQ, double s, double c, int nrows = HPGB_ATANQ_ROWS_POLAR) {
    double hi, lo;
    hpgb_atan2_tab_dd(TQ, s, c, hi, lo, nrows);
    return hi + lo;
}

#endif  // HPGB_MATH_H
