// hpgb_fp.h -- floating-point per-element cores: angle <-> pixel, vector <-> pixel,
// interpolation weights, neighbours.  `__host__ __device__` like hpgb_core.h.
//
// Parity rules: every +,-,*,/,sqrt,fmod below is a single IEEE operation in the
// reference's order (the translation units are compiled with implicit FMA contraction off);
// cos/sin/acos/atan2 come from hpgb_math.h (correctly rounded in practice, which is what the
// reference's glibc returns for ~99.9 % of arguments).
#ifndef HPGB_FP_H
#define HPGB_FP_H

#include "hpgb_core.h"
#include "hpgb_math.h"
#include "hpgb_reorder.h"

// same decimal literals as hpgeom/healpix_geom.h:31-36
#define HPGB_PI 3.141592653589793238462643383279502884197
#define HPGB_TWOPI 6.283185307179586476925286766559005768394
#define HPGB_HALFPI 1.570796326794896619231321691639751442099
#define HPGB_INV_HALFPI 0.6366197723675813430755350534900574

// double -> int64 the way C's cast does for in-range values (truncation)
HPGB_HD int64_t hpgb_d2ll(double v) {
#if defined(__CUDA_ARCH__)
    return __double2ll_rz(v);
#else
    return (int64_t)v;
#endif
}

// hpgeom/healpix_geom.c:112-117
HPGB_HD double hpgb_fmodulo(double v1, double v2) {
    if (v1 >= 0) return (v1 < v2) ? v1 : fmod(v1, v2);
    const double t = fmod(v1, v2) + v2;
    return (t == v2) ? 0. : t;
}

// a / 180.0, bit-identical to the IEEE quotient: RN(1/180) and two residual corrections (checked against
// operator / on 1.5e9 values of the input domain and random bit patterns down to 2^-900; zeros, the
// subnormal range and NaN take the division itself).  The reference's degree conversion is
// (x * pi) / 180.0 (hpgeom/hpgeom_utils.c:124-125); div.rn.f64 costs ~25 instructions, this 5.
HPGB_HD double hpgb_div180(double a) {
#if defined(__CUDA_ARCH__)
    if (!(fabs(a) > 1e-270)) return a / 180.0;
    const double y = 0x1.6c16c16c16c17p-8;  // RN(1/180)
    double q = a * y;
    q = fma(fma(-180.0, q, a), y, q);
    q = fma(fma(-180.0, q, a), y, q);
    return q;
#else
    return a / 180.0;
#endif
}

// Input angles -> (theta, phi), with the reference's range checks.
// hpgeom/hpgeom_utils.c:119-143 (lonlat) and :48-60 (theta/phi).  Returns true when valid.
template <bool LONLAT, bool DEGREES>
HPGB_HD bool hpgb_angles_exact(double a, double b, double &theta, double &phi) {
    if (LONLAT) {
        double lon = a, lat = b;
        if (DEGREES) {
            lon = hpgb_fmodulo(hpgb_div180(lon * HPGB_PI), HPGB_TWOPI);
            lat = hpgb_div180(lat * HPGB_PI);
        } else {
            lon = hpgb_fmodulo(lon, HPGB_TWOPI);
        }
        if (lat < -HPGB_HALFPI || lat > HPGB_HALFPI) return false;
        phi = lon;
        theta = -lat + HPGB_HALFPI;
        return true;
    }
    theta = a;
    phi = b;
    if (a < 0.0 || a > HPGB_PI) return false;
    if (b < 0 || b > HPGB_TWOPI) return false;
    return true;
}

// (z, phi, sin theta) -> pixel: hpgeom/healpix_geom.c:229-302, operation for operation.
template <bool NEST>
HPGB_HD int64_t hpgb_loc2pix_exact(const hpgb_geom &g, double z, double phi, double sth, bool have_sth) {
    const double za = fabs(z);
    const double tt = hpgb_fmodulo(phi * HPGB_INV_HALFPI, 4.0);
    const int64_t ns = g.nside;
    const double dns = (double)ns;
    if (!NEST) {
        if (za <= 2.0 / 3.0) {
            const int64_t nl4 = 4 * ns;
            const double t1 = dns * (0.5 + tt);
            const double t2 = dns * z * 0.75;
            const int64_t jp = hpgb_d2ll(t1 - t2);
            const int64_t jm = hpgb_d2ll(t1 + t2);
            const int64_t ir = ns + 1 + jp - jm;
            const int64_t kshift = 1 - (ir & 1);
            const int64_t u = jp + jm - ns + kshift + 1 + nl4 + nl4;
            const int64_t ip = (g.order > 0) ? ((u >> 1) & (nl4 - 1)) : ((u >> 1) % nl4);
            return g.ncap + (ir - 1) * nl4 + ip;
        }
        const double tp = tt - (double)hpgb_d2ll(tt);
        const double tmp = ((za < 0.99) || (!have_sth)) ? dns * sqrt(3 * (1 - za)) : dns * sth / sqrt((1. + za) / 3.);
        const int64_t jp = hpgb_d2ll(tp * tmp);
        const int64_t jm = hpgb_d2ll((1.0 - tp) * tmp);
        const int64_t ir = jp + jm + 1;
        const int64_t ip = hpgb_d2ll(tt * (double)ir);
        return (z > 0.) ? 2 * ir * (ir - 1) + ip : g.npix - 2 * ir * (ir + 1) + ip;
    }
    if (za <= 2.0 / 3.0) {
        const double t1 = dns * (0.5 + tt);
        const double t2 = dns * (z * 0.75);
        const int64_t jp = hpgb_d2ll(t1 - t2);
        const int64_t jm = hpgb_d2ll(t1 + t2);
        const int64_t ifp = jp >> g.order;
        const int64_t ifm = jm >> g.order;
        const uint32_t face = (uint32_t)((ifp == ifm) ? (ifp | 4) : ((ifp < ifm) ? ifp : (ifm + 8)));
        const uint32_t ix = (uint32_t)(jm & (ns - 1));
        const uint32_t iy = (uint32_t)(ns - (jp & (ns - 1)) - 1);
        return hpgb_xyf2nest(g.order, ix, iy, face);
    }
    int ntt = (int)tt;
    if (ntt >= 4) ntt = 3;
    const double tp = tt - ntt;
    const double tmp = ((za < 0.99) || (!have_sth)) ? dns * sqrt(3 * (1 - za)) : dns * sth / sqrt((1. + za) / 3.);
    int64_t jp = hpgb_d2ll(tp * tmp);
    int64_t jm = hpgb_d2ll((1.0 - tp) * tmp);
    if (jp >= ns) jp = ns - 1;
    if (jm >= ns) jm = ns - 1;
    return (z >= 0) ? hpgb_xyf2nest(g.order, (uint32_t)(ns - jm - 1), (uint32_t)(ns - jp - 1), (uint32_t)ntt)
                    : hpgb_xyf2nest(g.order, (uint32_t)jp, (uint32_t)jm, (uint32_t)(ntt + 8));
}

// hpgeom/healpix_geom.c:153-159 with the accurate cos/sin
template <bool NEST>
HPGB_HD int64_t hpgb_ang2pix_exact(const hpgb_geom &g, const double *T, double theta, double phi) {
    double s, c;
    hpgb_sincos_cr(T, theta, s, c);
    if ((theta < 0.01) || (theta > 3.14159 - 0.01)) return hpgb_loc2pix_exact<NEST>(g, c, phi, 0.0, false);
    return hpgb_loc2pix_exact<NEST>(g, c, phi, s, true);
}

// Full reference semantics for one (a, b): checks + conversion.  Returns false for a bad element.
template <bool NEST, bool LONLAT, bool DEGREES>
HPGB_HD_NOINLINE bool hpgb_angle_to_pixel_exact(const hpgb_geom &g, const double *T, double a, double b,
                                                int64_t &pix) {
    double theta, phi;
    if (!hpgb_angles_exact<LONLAT, DEGREES>(a, b, theta, phi)) {
        pix = -1;
        return false;
    }
    pix = hpgb_ang2pix_exact<NEST>(g, T, theta, phi);
    return true;
}

// ---------------------------------------------------------------------------------------------
// angle_to_pixel, fast path (NEST and, for power-of-two nside, RING).
//
// The pixel index is a step function of the angles; away from the steps it does not care about
// the last bits of cos(theta).  The fast path therefore uses cheap, branch-free fp64 (~35 fp64
// pipe operations per point: no division, no sqrt, no full-range cos/sin) with ~1e-15 relative
// error, keeps the two edge-line coordinates in fixed point, and hands the point to the exact
// path above only when a coordinate is within `margin` of an integer (a pixel edge) or a
// discrete decision (equatorial/polar split, face boundary, range check) is closer than its
// guard band.  margin = nside * 2^-44 covers the fast path's error AND the reference's own
// rounding noise (its sqrt(3(1-|z|)) amplifies a half-ulp error of cos up to 2^-48 relative),
// so fast result == exact-path result whenever the fast result is kept.  The exact path is
// taken for ~1e-4 of uniformly distributed points at nside 2^29 (less at smaller nside).
//
// Within 0.01 rad of a pole the reference itself is noisy: it skips sin(theta) there
// (hpgeom/healpix_geom.c:154-157) and falls back on nside*sqrt(3(1-|z|)), where the rounding of
// z = cos(theta) alone moves the result by up to nside * 7e-17 / theta -- more than the margin for
// theta < ~5e-3.  Those points (1e-4 of the sphere) always take the exact path, which reproduces
// the reference's rounding bit for bit.
//
// shortcuts ("z and phi shortcuts" of the design):
//   equatorial |z| <= 2/3 : z = sin(lat)                    (odd polynomial, |lat| <= 0.73)
//   polar caps            : nside*sqrt(3(1-|z|)) = nside*sqrt(6)*sin(c/2), c = colatitude from the
//                           nearest pole (same polynomial, argument <= 0.42; no cancellation)
//   tt = phi*2/pi mod 4   : lon/90 mod 4 directly in degrees
// ---------------------------------------------------------------------------------------------
#define HPGB_ASIN_2_3 0.72972765622696636345479665981332  // asin(2/3)

// v with its sign flipped when s is negative (one LOP3 on the high word)
HPGB_HD double hpgb_copysign_xor(double v, double s) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(v) ^ (__double2hiint(s) & (int)0x80000000), __double2loint(v));
#else
    return signbit(s) ? -v : v;
#endif
}

HPGB_HD uint32_t hpgb_hi32(double v) {
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

HPGB_HD uint32_t hpgb_lo32(double v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(v);
#else
    union {
        double d;
        uint64_t u;
    } c;
    c.d = v;
    return (uint32_t)c.u;
#endif
}
// low 32 bits of (hi:lo) >> s, 0 < s < 32 (SHF.R.U64 on the device)
HPGB_HD uint32_t hpgb_funnel_r(uint32_t lo, uint32_t hi, int s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return (lo >> s) | (hi << (32 - s));
#endif
}

// Per-nside constants of the fast path, host-computed and passed by value with the geometry.
// The polynomial coefficients and unit factors travel here too: kernel parameters live in the
// constant bank, so fp64 instructions read them as operands instead of materialising 64-bit
// immediates (two extra instructions per constant per element otherwise).  Everything that
// multiplies an edge-line coordinate is pre-scaled by 2^F (exact), so the coordinates come out
// directly in fixed point.
struct hpgb_a2p_fast {
    double ns_s, ns_half_s, ns_075_s, ns_sqrt6_s;  // (nside, nside/2, 0.75 nside, sqrt(6) nside) * 2^F
    double c3, c5, c7, c9, c11, c13, c15;          // sin(a) = a + a^3 (c3 + c5 a^2 + ...), |a| <= 0.74
    double k_lat, k_hcth, k_lon;  // radians per latitude unit; half of it; tt per longitude unit
    double pole, pole_lo;         // input latitude of the pole (90, or pi/2 as head + tail)
    double lat_max, lon_max;      // fast-path range guards
    double asin23;                // asin(2/3): equatorial / polar split in latitude
    double d_ns_m;                // (0.75 - sqrt(6)) nside * 2^F: equatorial minus polar multiplier of the sine
    uint32_t fbits, fmask, margin, enabled;
    uint32_t mask_hi, pad2_;      // fraction bits at or above 2 * margin
    uint32_t face_shift_hi;       // 2*order - 32 when order >= 16 (face bits live in the high word)
    uint32_t ord;                 // order, 0 when nside is not a power of two (fast path disabled then)
    double ns_s_magic;            // ns_s + 2^52 (HPGB_A2P_MAGIC: the edge-line coordinates are read off the mantissa)
    uint32_t margin2, pad3_;      // 2 * margin (HPGB_A2P_MAGIC: any value, not only powers of two)
};

template <bool LONLAT, bool DEGREES>
HPGB_HD hpgb_a2p_fast hpgb_make_a2p_fast(const hpgb_geom &g) {
    hpgb_a2p_fast f;
    const double dns = (double)g.nside;
    const int order = g.order < 0 ? 0 : g.order;
#if !defined(HPGB_A2P_V1)
    // F = 20 fraction bits for every order: coordinate * 2^20 < 6 * 2^29 * 2^20 < 2^52, so  value + 2^52  lies in
    // [2^52, 2^53) and its mantissa IS the coordinate in fixed point, rounded to the nearest unit (no F2I).  margin =
    // nside * 2^-44 in these units (32 at order 29), never below 16 units (2^-16 pixel: the fast path's error is a few
    // 2^-52 of a value < 2^32), plus 2 units for the two roundings to the unit grid (S + 2^52, then +- D).
    f.fbits = 20u;
    f.fmask = 0xFFFFFu;
    f.margin = (order >= 25 ? (1u << (order - 24)) : 1u);
    f.margin = (f.margin < 16u ? 16u : f.margin) + 2u;
#else
    // F fraction bits, always inside the low 32-bit word; margin = nside * 2^-44 in fixed-point
    // units (never below 16 units), see the derivation above hpgb_angle_to_pixel_nest_fast
    f.fbits = order >= 27 ? (uint32_t)(58 - order) : 32u;
    f.fmask = f.fbits >= 32 ? 0xFFFFFFFFu : ((1u << f.fbits) - 1u);
    f.margin = order >= 26 ? (1u << 14) : (order >= 16 ? (1u << (order - 12)) : 16u);
#endif
    const double scale = ldexp(1.0, (int)f.fbits);
    f.ns_s = dns * scale;
    f.ns_half_s = 0.5 * dns * scale;
    f.ns_075_s = 0.75 * dns * scale;
    f.ns_sqrt6_s = 2.4494897427831780981972840747059 * dns * scale;
    f.d_ns_m = f.ns_075_s - f.ns_sqrt6_s;
#if defined(HPGB_SIN_TAYLOR15)
    f.c3 = -1.6666666666666666667e-1;
    f.c5 = 8.3333333333333333333e-3;
    f.c7 = -1.9841269841269841270e-4;
    f.c9 = 2.7557319223985890653e-6;
    f.c11 = -2.5052108385441718775e-8;
    f.c13 = 1.6059043836821614599e-10;
    f.c15 = -7.6471637318198164759e-13;
#else
    // degree 13, interpolated at the Chebyshev nodes of a^2 in [0, 0.74^2] (mpmath, 200 bits): max relative error
    // 2^-57.3 over |a| <= 0.74 -- what the Taylor coefficients need degree 15 for; one DFMA (two issue cycles) less
    f.c3 = -0x1.5555555555555p-3;
    f.c5 = 0x1.1111111110e1bp-7;
    f.c7 = -0x1.a01a019f1d752p-13;
    f.c9 = 0x1.71de3869568b0p-19;
    f.c11 = -0x1.ae60f378ed40dp-26;
    f.c13 = 0x1.5e63d28acf0d3p-33;
    f.c15 = 0.0;
#endif
    if (LONLAT && DEGREES) {
        f.k_lat = 1.7453292519943295769e-2;   // pi/180
        f.k_hcth = 8.7266462599716478846e-3;  // pi/360
        f.k_lon = 1.1111111111111111111e-2;   // 1/90
        f.pole = 90.0;
        f.pole_lo = 0.0;
        f.lat_max = 89.42;  // see below: colatitude < 0.0101 rad goes to the exact path
        f.lon_max = 720.0;
    } else {
        f.k_lat = 1.0;
        f.k_hcth = 0.5;
        f.k_lon = HPGB_INV_HALFPI;
        f.pole = HPGB_PIO2_1;
        f.pole_lo = HPGB_PIO2_2;
        f.lat_max = 1.5606;
        f.lon_max = 12.5;
    }
    f.asin23 = 0.72972765622696636345479665981332;
    f.enabled = (g.order >= 0) ? 1u : 0u;
    f.face_shift_hi = order >= 16 ? (uint32_t)(2 * order - 32) : 0u;
    f.ord = (uint32_t)order;
    f.mask_hi = f.fmask & ~(2u * f.margin - 1u);
    f.pad2_ = 0;
    f.ns_s_magic = f.ns_s + 4503599627370496.0;  // exact: ns_s is an integer < 2^50
    f.margin2 = 2u * f.margin;
    f.pad3_ = 0;
    return f;
}

HPGB_HD double hpgb_sin_poly(const hpgb_a2p_fast &f, double a) {  // |a| <= 0.74, error < 1 ulp
    const double a2 = a * a;
#if defined(HPGB_SIN_TAYLOR15)
    double p = fma(a2, f.c15, f.c13);
    p = fma(p, a2, f.c11);
#else
    double p = fma(a2, f.c13, f.c11);
#endif
    p = fma(p, a2, f.c9);
    p = fma(p, a2, f.c7);
    p = fma(p, a2, f.c5);
    p = fma(p, a2, f.c3);
    return fma(a * a2, p, a);
}

// Shared tail of the NEST fast paths (angle_to_pixel, vector_to_pixel): from floor(tt), tp - 1/2,
// the zone flag e (1.0 equatorial / 0.0 polar), v = (3/4) nside z (equatorial) or w (polar), both
// scaled by 2^F, and the hemisphere (sign of `hemi`) to the pixel.  Returns the margin guard bit.
// NEST: Morton interleave (HI: nside >= 2^16, face bits in the high word only).  RING: the same
// (face, ix, iy) goes through xyf2ring -- equal to the reference's direct ring formulas
// (hpgeom/healpix_geom.c:236-262): the equatorial branch uses the same jp, jm; in the caps the
// reference takes ip = floor(tt * ir), which equals ntt * ir + jp whenever tp * tmp and
// (1 - tp) * tmp are at least `margin` away from integers (a convex-combination argument:
// frac(tp * ir) lies between frac(tp * tmp) and 1 - frac((1 - tp) * tmp)), i.e. whenever the
// fast result is kept.  rk is only read for RING.
#if defined(HPGB_A2P_V1)
template <bool NEST, bool HI>
HPGB_HD unsigned hpgb_sd_to_pix(const hpgb_geom &g, const hpgb_a2p_fast &f, const hpgb_rk &rk, double nttd, double tq,
                                double e, double v, double hemi, int64_t &pix) {
    // ---- edge-line coordinates, already scaled to fixed point with F fraction bits ----
    const double W = fma(e, f.ns_s - v, v);
    const double Dp = hpgb_copysign_xor(fma(0.5, v, -f.ns_s), hemi);  // +-(w/2 - nside)
    const double D = fma(e, -v - Dp, Dp);
#if 0
    // S carries 2^52: S +- D lands in [2^52, 2^53), where one ulp is one fixed-point unit -- the sum's mantissa is
    // the coordinate (20 fraction bits), its integer part one funnel shift away.  The exponent field (0x433) sits
    // above bit 51 and falls off the 32-bit result of the shift.
    const double S = fma(tq, W, fma(nttd, f.ns_s, f.ns_s_magic));
    const double r1 = S + D, r2 = S - D;
    const uint32_t l1 = hpgb_lo32(r1), l2 = hpgb_lo32(r2);
    const unsigned guard = (unsigned)(((l1 + f.margin) & 0xFFFFFu) < f.margin2) | (unsigned)(((l2 + f.margin) & 0xFFFFFu) < f.margin2);
    const uint32_t j1 = hpgb_funnel_r(l1, hpgb_hi32(r1), 20), j2 = hpgb_funnel_r(l2, hpgb_hi32(r2), 20);
#else
    const double S = fma(tq, W, fma(nttd, f.ns_s, f.ns_s));
    const uint64_t x1 = (uint64_t)hpgb_d2ll(S + D);
    const uint64_t x2 = (uint64_t)hpgb_d2ll(S - D);
    // within `margin` of an integer?  (the fraction sits in the low word; 2 * margin is a power of two)
    const unsigned guard = (unsigned)((((uint32_t)x1 + f.margin) & f.mask_hi) == 0u) | (unsigned)((((uint32_t)x2 + f.margin) & f.mask_hi) == 0u);
    const uint32_t j1 = (uint32_t)(x1 >> f.fbits), j2 = (uint32_t)(x2 >> f.fbits);
#endif
    const uint32_t ifp = j1 >> f.ord, ifm = j2 >> f.ord;
    if (NEST && flut) {
        // j1, j2 ARE the edge-line coordinates jp, jm of ring_to_nest's belt pass (hpgb_reorder.h, "folded coordinates"):
        // the 16-entry table indexed by (ifp, ifm) supplies the two constants of X = jm + lx, Y = ly - jp, whose Morton
        // interleave is the pixel number with the face bits already in place -- one load and two adds instead of the face
        // selection, the two masks and the face shift-add.  The index is masked: an untrusted point (guard set, result
        // discarded) may carry any j1, j2 and must not read outside the table.
        const hpgb_u32x2 fl = flut[(2u * ifp + ifm + 1u) & 15u];
        pix = (int64_t)hpgb_interleave(j2 + fl.x, fl.y - j1);
        return guard;
    }
    const uint32_t nsm1 = (uint32_t)g.nside - 1u;
    const uint32_t face = hpgb_sel(ifp == ifm, ifp | 4u, hpgb_sel(ifp < ifm, ifp, ifm + 8u));
    const uint32_t ix = j2 & nsm1;
    const uint32_t iy = ~j1 & nsm1;  // nside - 1 - (j1 mod nside)
    if (!NEST) {
        pix = hpgb_xyf2ring_k(rk, ix, iy, face);
    } else if (HI) {
        const uint64_t m = hpgb_interleave(ix, iy);
        pix = (int64_t)(m + ((uint64_t)(face << f.face_shift_hi) << 32));
    } else {
        pix = hpgb_xyf2nest((int)f.ord, ix, iy, face);
    }
    return guard;
}

#else
// The fp64 pipe of sm_100a does not run beside the integer pipes: a DFMA / DADD / DSETP costs the warp scheduler
// two issue cycles, any 32-bit instruction one, and nothing overlaps them (tools/ubench/pipes2.cu: 2 DFMA + 2 LOP3
// = 0.71 instructions / clock / scheduler, the angle_to_pixel mix 0.73 -- the kernel runs at 0.67).  So the tail
// below spends 32-bit selects, not fp64 blends, on the equatorial / polar split, never forms floor() or a
// float -> int conversion (XU pipe, 9 cycles a warp), and adds the quarter-turn count in the integer domain:
//   ntt : quarter turns of the longitude, floor(tt) mod 4;   tq = frac(tt) - 1/2;   eq : |z| <= 2/3
//   v   : (3/4) nside z (equatorial) or w = nside sqrt(3 (1 - |z|)) (polar), scaled by 2^20
//   x1 = S + D, x2 = S - D,  S = nside + tq W + 2^52  (W = nside / w),  D = -v (equatorial) or +-(w/2 - nside)
// S +- D lands in [2^52, 2^53), where one ulp is one fixed-point unit: the sum's mantissa IS the coordinate (20
// fraction bits, rounded to nearest), its integer part one funnel shift away -- the exponent field sits above
// bit 51 and falls off the 32-bit result of the shift.  ntt * nside is added to both integer parts afterwards.
// entry e = 8 ntt + i of the 32-entry folded-face table of the NEST fast paths (see hpgb_sd_to_pix): i = 2 (jf1 >> k) +
// (jf2 >> k) + 1 with the coordinates taken WITHOUT their quarter turn (jf1 >> k in {0, 1}, jf2 >> k - jf1 >> k in {-1, 0, 1}:
// i = 1 .. 5); the entry is ring_to_nest's (lx, ly) for (jp >> k, jm >> k) = (jf1 >> k + ntt, jf2 >> k + ntt) with the turn
// ntt nside added to lx and taken from ly
HPGB_HD void hpgb_a2p_fold_entry(const hpgb_rk &c, uint32_t e, uint32_t &x, uint32_t &y) {
    const uint32_t t = e >> 3, i = e & 7u;
    x = y = 0u;
    if (i < 1u || i > 5u) return;
    uint32_t lx, ly;
    hpgb_r2n_lut_entry(c, i + 3u * t, lx, ly);
    x = lx + (t << c.k);
    y = ly - (t << c.k);
}
template <bool NEST, bool HI>
HPGB_HD unsigned hpgb_sd_to_pix(const hpgb_geom &g, const hpgb_a2p_fast &f, const hpgb_rk &rk, uint32_t ntt, double tq,
                                bool eq, double v, double hemi, int64_t &pix, const hpgb_u32x2 *flut = nullptr) {
    const double W = eq ? f.ns_s : v;
    const double S = fma(tq, W, f.ns_s_magic);
    const double Dp = hpgb_copysign_xor(fma(0.5, v, -f.ns_s), hemi);  // +-(w/2 - nside)
    const double D = eq ? -v : Dp;
    const double r1 = S + D, r2 = S - D;
    const uint32_t l1 = hpgb_lo32(r1), l2 = hpgb_lo32(r2);
    // within `margin` units of an integer?
    const unsigned guard = (unsigned)(((l1 + f.margin) & 0xFFFFFu) < f.margin2) | (unsigned)(((l2 + f.margin) & 0xFFFFFu) < f.margin2);
    const uint32_t jf1 = hpgb_funnel_r(l1, hpgb_hi32(r1), 20), jf2 = hpgb_funnel_r(l2, hpgb_hi32(r2), 20);
    if (NEST && flut) {
        // jf1 + ntt nside, jf2 + ntt nside ARE the edge-line coordinates jp, jm of ring_to_nest's belt pass (hpgb_reorder.h,
        // "folded coordinates"): a table indexed by (jp >> k, jm >> k) supplies the two constants of X = jm + lx,
        // Y = ly - jp, whose Morton interleave is the pixel number with the face bits already in place -- one load and two
        // adds instead of the face selection, the two masks and the face shift-add.  The quarter turn rides in the table
        // as well (hpgb_a2p_fold_entry: 4 x 8 entries, lx + ntt nside, ly - ntt nside), so it is never added to the
        // coordinates, and the index is masked instead of ntt: an untrusted point (guard set, result discarded) may carry
        // any coordinates and must not read outside the table.
        const hpgb_u32x2 fl = flut[(8u * ntt + 2u * (jf1 >> f.ord) + (jf2 >> f.ord) + 1u) & 31u];
        pix = (int64_t)hpgb_interleave(jf2 + fl.x, fl.y - jf1);
        return guard;
    }
    const uint32_t turn = (ntt & 3u) << f.ord;  // ntt * nside
    const uint32_t j1 = jf1 + turn, j2 = jf2 + turn;
    const uint32_t ifp = j1 >> f.ord, ifm = j2 >> f.ord;
    const uint32_t nsm1 = (uint32_t)g.nside - 1u;
    const uint32_t face = hpgb_sel(ifp == ifm, ifp | 4u, hpgb_sel(ifp < ifm, ifp, ifm + 8u));
    const uint32_t ix = j2 & nsm1;
    const uint32_t iy = ~j1 & nsm1;  // nside - 1 - (j1 mod nside)
    if (!NEST) {
        pix = hpgb_xyf2ring_k(rk, ix, iy, face);
    } else if (HI) {
        const uint64_t m = hpgb_interleave(ix, iy);
        pix = (int64_t)(m + ((uint64_t)(face << f.face_shift_hi) << 32));
    } else {
        pix = hpgb_xyf2nest((int)f.ord, ix, iy, face);
    }
    return guard;
}
#endif

// Fast path for one point: candidate pixel + whether it can be trusted (false -> the caller runs
// the exact path).  Branch-free apart from the equatorial/polar integer tail; the guard
// conditions are accumulated bitwise so no short-circuit branches appear in the hot loop.
// HI: nside >= 2^16, i.e. the face number only touches the high word of the pixel.
#if defined(HPGB_A2P_V1)
template <bool NEST, bool LONLAT, bool DEGREES, bool HI>
HPGB_HD uint32_t hpgb_angle_to_pixel_fast_g(const hpgb_geom &g, const hpgb_a2p_fast &f, const hpgb_rk &rk, double a,
                                            double b, int64_t &pix) {
    // ---- latitude (radians, signed), half colatitude from the nearest pole, tt in [0,4) ----
    double lat, hcth, tt;
    unsigned guard;
    if (LONLAT) {
        lat = DEGREES ? b * f.k_lat : b;
        hcth = DEGREES ? (f.pole - fabs(b)) * f.k_hcth : ((f.pole - fabs(b)) + f.pole_lo) * 0.5;
        const double t = a * f.k_lon;  // lon/90 or lon*2/pi
        tt = fma(-4.0, floor(t * 0.25), t);
        guard = (unsigned)!(fabs(b) < f.lat_max) | (unsigned)!(fabs(a) <= f.lon_max);
    } else {
        lat = (f.pole - a) + f.pole_lo;  // pi/2 - theta, accurate also near the equator
        hcth = 0.5 * ((a < f.pole) ? a : (HPGB_PI_1 - a) + HPGB_PI_2);
        tt = b * f.k_lon;                // the reference's own expression
        guard = (unsigned)!(a > 0.0102) | (unsigned)!(a < 3.1313) | (unsigned)!(b >= 0.0) | (unsigned)!(b < 6.2831853);
    }
    const double al = fabs(lat);
    const bool eq = al <= f.asin23;
    guard |= (unsigned)(fabs(al - f.asin23) < 1e-11);
    // The integer ALU pipe is what bounds this kernel, so the equatorial / polar split is resolved
    // in fp64 arithmetic (a blend with e = 1.0 / 0.0) instead of integer selects, and BOTH zones are
    // brought to the same pair of edge-line coordinates, so one integer tail serves both:
    //   x1 = S + D, x2 = S - D,   S = nside (ntt + 1) + (tp - 1/2) W
    //     equatorial : W = nside,  D = -(3/4) nside sin(lat)           -> nside (1/2 + tt) -+ (3/4) nside z
    //     polar      : W = w,      D = +-(w/2 - nside)  (+ north)      ,  w = sqrt(6) nside sin(colat/2)
    //                  north: x1 = nside ntt + tp w,  x2 = nside (ntt + 2) - (1 - tp) w;  south: swapped
    //   For a polar point floor(x1), floor(x2) are  nside ntt + jp  and  nside (ntt + 1) + (nside-1-jm)
    //   (jp, jm as in the reference, hpgeom/healpix_geom.c:289-292), so the equatorial rule "face
    //   from (x1 >> k, x2 >> k), ix = x2 mod nside, iy = nside-1 - x1 mod nside" yields the
    //   reference's polar (face, ix, iy) too: north face = ntt, (nside-1-jm, nside-1-jp); south
    //   face = ntt + 8, (jp, jm).  The reference's clamps jp, jm <= nside-1 only act at |z| = 2/3
    //   exactly, inside the guard band above.
    const double e = eq ? 1.0 : 0.0;
    const double nttd = floor(tt);
    const double tq = (tt - nttd) - 0.5;  // tp - 1/2
    guard |= (unsigned)!(fabs(tq) < 0.5 - 3.8146972656250000e-06);  // tt within 2^-18 of an integer (face edge / wrap)
    // ---- one sine for both zones ----
    const double s = hpgb_sin_poly(f, fma(e, lat - hcth, hcth));
    const double v = fma(e, f.d_ns_m, f.ns_sqrt6_s) * s;  // (3/4) nside z   or   w
    guard |= hpgb_sd_to_pix<NEST, HI>(g, f, rk, nttd, tq, e, v, lat, pix);
    return guard | (f.enabled ^ 1u);
}
#else
// 1.5 * 2^52: adding it rounds a double of magnitude < 2^51 to the nearest integer, which then sits in the low
// mantissa bits (two's complement for negative values)
#define HPGB_RND_MAGIC 6755399441055744.0
template <bool NEST, bool LONLAT, bool DEGREES, bool HI>
HPGB_HD uint32_t hpgb_angle_to_pixel_fast_g(const hpgb_geom &g, const hpgb_a2p_fast &f, const hpgb_rk &rk, double a,
                                            double b, int64_t &pix, const hpgb_u32x2 *flut = nullptr) {
    // ---- latitude (radians, signed), half colatitude from the nearest pole, tt - 1/2 ----
    double lat, hcth, u;
    unsigned guard;
    if (LONLAT) {
        lat = DEGREES ? b * f.k_lat : b;
        hcth = DEGREES ? (f.pole - fabs(b)) * f.k_hcth : ((f.pole - fabs(b)) + f.pole_lo) * 0.5;
        u = fma(a, f.k_lon, -0.5);  // lon/90 - 1/2 or lon*2/pi - 1/2: one rounding
        guard = (unsigned)!(fabs(b) < f.lat_max) | (unsigned)!(fabs(a) <= f.lon_max);
    } else {
        lat = (f.pole - a) + f.pole_lo;  // pi/2 - theta, accurate also near the equator
        hcth = 0.5 * ((a < f.pole) ? a : (HPGB_PI_1 - a) + HPGB_PI_2);
        u = fma(b, f.k_lon, -0.5);
        guard = (unsigned)!(a > 0.0102) | (unsigned)!(a < 3.1313) | (unsigned)!(b >= 0.0) | (unsigned)!(b < 6.2831853);
    }
    // equatorial / polar split on the HIGH WORD of |lat| (one LOP3 + two 32-bit compares instead of two fp64 compares
    // and a subtraction, each of which costs two issue cycles): every latitude whose high word is that of asin(2/3) or
    // one of its two neighbours -- within 2^-20 relative of the split, 1.5e-6 of the sphere -- goes to the exact path,
    // and outside that band the high words alone order |lat| and asin(2/3) correctly
    const uint32_t alh = hpgb_hi32(lat) & 0x7FFFFFFFu, a23h = hpgb_hi32(f.asin23);
    const bool eq = alh < a23h;
    guard |= (unsigned)(alh - (a23h - 1u) < 3u);
    // floor(tt) = RN(tt - 1/2) unless tt is (nearly) an integer -- and then the point is on a face edge / the wrap
    // and goes to the exact path anyway (|tq| guard).  The rounding is one addition of 1.5 * 2^52; the integer is
    // read off the sum's low word, so a negative or wrapped longitude needs nothing extra: ntt = floor(tt) mod 4.
    const double r = u + HPGB_RND_MAGIC;
    const double tq = u - (r - HPGB_RND_MAGIC);  // frac(tt) - 1/2, exact
    const uint32_t ntt = hpgb_lo32(r);  // floor(tt) modulo 2^32: hpgb_sd_to_pix takes it modulo 4
    guard |= (unsigned)!(fabs(tq) < 0.5 - 3.8146972656250000e-06);  // tt within 2^-18 of an integer
    // ---- one sine for both zones: z = sin(lat) (equatorial), w = sqrt(6) nside sin(colat / 2) (polar) ----
    const double s = hpgb_sin_poly(f, eq ? lat : hcth);
    const double v = (eq ? f.ns_075_s : f.ns_sqrt6_s) * s;  // (3/4) nside z   or   w
    guard |= hpgb_sd_to_pix<NEST, HI>(g, f, rk, ntt, tq, eq, v, lat, pix, flut);
    return guard | (f.enabled ^ 1u);
}
#endif


// ---------------------------------------------------------------------------------------------
// vector_to_pixel, fast path (same contract as the angle_to_pixel one: candidate pixel + a
// guard word, 0 = trusted; anything else re-runs on hpgb_vec2pix, the reference's operation order).
//   z/|v| and sin^2(theta) from ONE reciprocal square root (no IEEE sqrt + division);
//   polar zone: 1 - |z| = sin^2 / (1 + |z|), no cancellation at any latitude;
//   phi: atan2 reduced to atan(r), |r| <= 1/31, with r = (num - c den) / (den + c num), c = k/16 picked
//   from an approximate quotient, 17-entry table of atan(k/16), odd polynomial of degree 11; the
//   quadrant only fixes floor(tt) and whether tp = u or 1 - u, so tt itself is never formed.
// All approximations are good to a few 2^-52, far inside the margin nside * 2^-44.
// T = the shared-memory table block (atan part at T + 4 * HPGB_SINCOS_ROWS).
// ---------------------------------------------------------------------------------------------
HPGB_HD double hpgb_rsqrt_fast(double x) {  // 1/sqrt(x), ~1 ulp, x normal and positive
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = 0.5 * x;
    y = fma(y, fma(-h * y, y, 0.5), y);
    y = fma(y, fma(-h * y, y, 0.5), y);
    return y;
#else
    return 1.0 / sqrt(x);
#endif
}
HPGB_HD double hpgb_rcp_fast(double x) {  // 1/x, ~1 ulp, x normal
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    return y;
#else
    return 1.0 / x;
#endif
}
HPGB_HD double hpgb_rcp_rough(double x) {  // ~2^-20 relative: only used to pick a table row
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    return 1.0 / x;
#endif
}

template <bool NEST, bool HI>
HPGB_HD uint32_t hpgb_vector_to_pixel_fast_g(const hpgb_geom &g, const hpgb_a2p_fast &f, const hpgb_rk &rk, const double *T,
                                             double x, double y, double z, int64_t &pix, const hpgb_u32x2 *flut = nullptr) {
    const double r2 = fma(y, y, x * x);
    const double len2 = fma(z, z, r2);
    // squares must stay normal numbers; |x| = |y| = 0 (phi = 0 by convention) goes to the exact path
    unsigned guard = (unsigned)!(len2 > 1e-280) | (unsigned)!(len2 < 1e280) | (unsigned)!(r2 > 1e-32 * len2);
    const double il = hpgb_rsqrt_fast(len2);
    const double nz = z * il;
    const double za = fabs(nz);
    const double s2 = (r2 * il) * il;  // sin^2(theta)
    const bool eq = za <= 2.0 / 3.0;
    guard |= (unsigned)(fabs(za - 2.0 / 3.0) < 1e-11);
    const double q = (3.0 * s2) * hpgb_rcp_fast(1.0 + za);  // 3 (1 - |z|)
    const double w = f.ns_s * (q * hpgb_rsqrt_fast(q));      // nside sqrt(3 (1 - |z|))
#if defined(HPGB_A2P_V1)
    const double e = eq ? 1.0 : 0.0;
    const double v = fma(e, f.ns_075_s * nz - w, w);
#else
    const double v = eq ? f.ns_075_s * nz : w;
#endif
    // ---- phi ----
    const double ax = fabs(x), ay = fabs(y);
    const bool swap = ay > ax;
    const double num = swap ? ax : ay, den = swap ? ay : ax;  // 0 <= num <= den, den > 0
#if defined(HPGB_V2P_RINT_V1)
    double fk = rint(16.0 * (num * hpgb_rcp_rough(den)));
    fk = fk >= 0.0 ? (fk <= 16.0 ? fk : 16.0) : 0.0;  // NaN / zero vectors (guarded above) must still index the table safely
    const int ik = (int)fk;
#else
    // table row = RN(16 num / den), read off a mantissa (hpgb_rint_i); only the INDEX is clamped -- a NaN / zero
    // vector is guarded above and merely has to address the table safely
    int ik;
    const double fk = hpgb_rint_i(16.0 * (num * hpgb_rcp_rough(den)), ik);
    ik = ik < 0 ? 0 : (ik > 16 ? 16 : ik);
#endif
    const double c = fk * 0.0625;
    const double r = fma(-c, den, num) * hpgb_rcp_fast(fma(c, num, den));
    const double rr = r * r;
    double p = fma(rr, -1.0 / 11.0, 1.0 / 9.0);
    p = fma(p, rr, -1.0 / 7.0);
    p = fma(p, rr, 1.0 / 5.0);
    p = fma(p, rr, -1.0 / 3.0);
    const double at = T[4 * HPGB_SINCOS_ROWS + ik] + fma(r * rr, p, r);  // atan(num/den) in [0, pi/4]
    const double u = at * HPGB_INV_HALFPI;                                    // in quarter turns, [0, 1/2]
    // quadrant: floor(tt) from the signs alone; tp = u or 1 - u by the parity of (swap, x < 0, y < 0)
    const uint32_t xs = hpgb_hi32(x) >> 31, ys = hpgb_hi32(y) >> 31;
    const bool flip = ((uint32_t)swap ^ xs ^ ys) != 0u;
    // tp - 1/2 = u - 1/2, or its exact negative 1/2 - u: one subtraction and a sign flip on the high word
    const double tq = hpgb_flip_sign_bit(u - 0.5, (uint32_t)flip << 31);
    guard |= (unsigned)!(fabs(tq) < 0.5 - 3.8146972656250000e-06);  // tt within 2^-18 of an integer (face edge / wrap)
#if defined(HPGB_A2P_V1)
    const double xn = xs ? 1.0 : 0.0;
    const double nttd = ys ? 3.0 - xn : xn;
    guard |= hpgb_sd_to_pix<NEST, HI>(g, f, rk, nttd, tq, e, v, nz, pix);
#else
    guard |= hpgb_sd_to_pix<NEST, HI>(g, f, rk, ys ? 3u - xs : xs, tq, eq, v, nz, pix, flut);
#endif
    return guard | (f.enabled ^ 1u);
}

// ---------------------------------------------------------------------------------------------
// pixel -> (z, phi, sin theta): hpgeom/healpix_geom.c:304-380, operation for operation
// (all IEEE-exact: integer work, int->double conversions, *, /, -, sqrt).
// ---------------------------------------------------------------------------------------------
template <bool NEST, bool POW2>
HPGB_HD void hpgb_pix2loc(const hpgb_geom &g, int64_t pix, double &z, double &phi, double &sth, bool &have_sth) {
    const int64_t ns = g.nside;
    have_sth = false;
    sth = 0.0;
    if (!NEST) {
        if (pix < g.ncap || pix >= g.npix - g.ncap) {  // polar caps
            const bool north = pix < g.ncap;
            const int64_t ip = north ? pix + 1 : g.npix - pix;
            const int64_t iring = (1 + hpgb_isqrt(2 * ip - 1)) >> 1;  // from the nearest pole
            const int64_t off = ip - 2 * iring * (iring - 1);
            const int64_t iphi = north ? off : 4 * iring + 1 - off;
            const double t = (double)(iring * iring) * g.fact2;
            z = north ? 1.0 - t : t - 1.0;
            if (north ? (z > 0.99) : (z < -0.99)) {
                sth = sqrt(t * (2. - t));
                have_sth = true;
            }
            phi = ((double)iphi - 0.5) * HPGB_HALFPI / (double)iring;
        } else {
            const int64_t nl4 = 4 * ns;
            const int64_t ip = pix - g.ncap;
            const int64_t t = POW2 ? (ip >> (g.order + 2)) : ip / nl4;
            const int64_t iring = t + ns, iphi = ip - nl4 * t + 1;
            const double fodd = ((iring + ns) & 1) ? 1 : 0.5;
            z = (double)(2 * ns - iring) * g.fact1;
            phi = ((double)iphi - fodd) * HPGB_PI * 0.75 * g.fact1;
        }
        return;
    }
    uint32_t ix, iy, face;
    hpgb_nest2xyf(g, pix, ix, iy, face);
    const int64_t jr = ((int64_t)hpgb_jrll(face) << g.order) - (int64_t)ix - (int64_t)iy - 1;
    int64_t nr;
    if (jr < ns) {
        nr = jr;
        const double t = (double)(nr * nr) * g.fact2;
        z = 1 - t;
        if (z > 0.99) {
            sth = sqrt(t * (2. - t));
            have_sth = true;
        }
    } else if (jr > 3 * ns) {
        nr = ns * 4 - jr;
        const double t = (double)(nr * nr) * g.fact2;
        z = t - 1;
        if (z < -0.99) {
            sth = sqrt(t * (2. - t));
            have_sth = true;
        }
    } else {
        nr = ns;
        z = (double)(2 * ns - jr) * g.fact1;
    }
    int64_t t = (int64_t)hpgb_jpll(face) * nr + (int64_t)ix - (int64_t)iy;
    if (t < 0) t += 8 * nr;
    phi = (nr == ns) ? 0.75 * HPGB_HALFPI * (double)t * g.fact1 : (0.5 * HPGB_HALFPI * (double)t) / (double)nr;
}

// Same results as hpgb_pix2loc for nside = 2^k, organised for the GPU: 32-bit integer work, one
// straight-line body for caps and belt (selects, no per-element branch, so the pixels a thread owns
// interleave), and sin(theta) left to the caller (only needed for |z| > 0.99, 1 % of the pixels):
// `tcap` is the reference's `tmp` = nr^2 * fact2 (sth = sqrt(tcap (2 - tcap))), `polar` tells
// whether the reference would have set have_sth.  Returns false for the ~1e-8 of RING cap pixels
// where the reference's uncorrected isqrt mis-rounds (the caller then uses hpgb_pix2loc).
// integer part of hpgb_pix2loc_k: zone flags, ring size / 4 (`nr`), the integer the reference converts for phi
// (`tphi`, minus `phi_off` in the RING scheme), 2 nside - ring for the belt (`zb`) and the ring number itself
// (`ring`, 1 .. 4 nside - 1 counted from the north pole)
template <bool NEST>
HPGB_HD bool hpgb_pix2zone_k(const hpgb_rk &c, int64_t pix, uint32_t &nr, uint32_t &tphi, int32_t &zb, bool &north,
                             bool &cap, double &phi_off, uint32_t &ring) {
    bool ok = true;
    phi_off = 0.0;  // RING: subtracted from the in-ring index before scaling
    if (NEST) {
        const uint32_t lo = (uint32_t)(uint64_t)pix, hi = (uint32_t)((uint64_t)pix >> 32);
        const uint32_t face = (uint32_t)((uint64_t)pix >> c.k2);
        uint32_t wl = lo & c.lmask, wh = hi & c.hmask;
        HPGB_DSWAP(wl, 0x22222222u, 1);
        HPGB_DSWAP(wh, 0x22222222u, 1);
        HPGB_DSWAP(wl, 0x0C0C0C0Cu, 2);
        HPGB_DSWAP(wh, 0x0C0C0C0Cu, 2);
        HPGB_DSWAP(wl, 0x00F000F0u, 4);
        HPGB_DSWAP(wh, 0x00F000F0u, 4);
        const uint32_t ix = hpgb_prmt(wl, wh, 0x6420), iy = hpgb_prmt(wl, wh, 0x7531);
        const uint32_t t = (face >> 2) * c.ns + c.nsm1 - ix - iy;  // ring - nside (two's complement)
        north = (int32_t)t < 0;
        cap = t > c.ns2;  // jr > 3 nside, or jr < nside (wraps)
        nr = hpgb_sel(cap, hpgb_sel(north, t + c.ns, c.ns3 - t), c.ns);
        zb = (int32_t)(c.ns - t);
        ring = t + c.ns;
        // jpll * nr + ix - iy, + 8 nr when negative (only face row 1, jpll = 0, can be negative)
        const uint32_t jp = 2u * (face & 3u) + 1u - ((face >> 2) & 1u);
        const uint32_t v = jp * nr + ix - iy;
        tphi = hpgb_sel(jp == 0u && ix < iy, v + 8u * nr, v);
    } else {
        uint32_t ipb_lo;
        const uint32_t tmp = hpgb_r2n_zone(c, (uint64_t)pix, ipb_lo);  // ring - nside
        north = (int32_t)tmp < 0;
        cap = hpgb_r2n_is_cap(c, tmp);
        // caps: the reference's isqrt (hpgeom/healpix_geom.c:60,307-310,326-329)
        const uint64_t up = (uint64_t)pix;
        const uint64_t ipm1 = north ? up : ((((uint64_t)c.npm1_hi << 32) | c.npm1_lo) - up);
#if defined(__CUDA_ARCH__)
        const uint32_t isq = __double2uint_rz(sqrt((double)(int64_t)(2 * ipm1 + 1) + 0.5));
#else
        const uint32_t isq = (uint32_t)(int64_t)sqrt((double)(int64_t)(2 * ipm1 + 1) + 0.5);
#endif
        const uint32_t ir = (isq + 1u) >> 1;
        const uint32_t q0 = (uint32_t)ipm1 - ir * (2u * ir - 2u);  // 0-based in-ring index counted from the pole side
        ok = !cap || q0 < 4u * ir;
        const uint32_t iphi_cap = hpgb_sel(north, q0 + 1u, 4u * ir - q0);
        nr = hpgb_sel(cap, ir, c.ns);
        zb = (int32_t)(c.ns - tmp);
        ring = hpgb_sel(cap, hpgb_sel(north, ir, c.ns4 - ir), tmp + c.ns);
        tphi = hpgb_sel(cap, iphi_cap, (ipb_lo & c.ns4m1) + 1u);
        phi_off = cap ? 0.5 : ((tmp & 1u) ? 1.0 : 0.5);  // belt: (iring + nside) & 1 ? 1 : 0.5, iring + nside = tmp + 2 nside
    }
    return ok;
}
// phi of a pixel centre from the integers above (hpgeom/healpix_geom.c:319,338,345 / :369-376)
template <bool NEST>
HPGB_HD double hpgb_zone_phi_k(const hpgb_geom &g, bool cap, uint32_t tphi, double phi_off, double dnr) {
    const double dt = hpgb_u2d(tphi);
    if (NEST) {
        const double pb = 0.75 * HPGB_HALFPI * dt * g.fact1;
        const double pc = hpgb_div_tame(0.5 * HPGB_HALFPI * dt, dnr);
        return cap ? pc : pb;
    }
    const double d = dt - phi_off;
    const double pb = d * HPGB_PI * 0.75 * g.fact1;
    const double pc = hpgb_div_tame(d * HPGB_HALFPI, dnr);
    return cap ? pc : pb;
}

template <bool NEST>
HPGB_HD bool hpgb_pix2loc_k(const hpgb_rk &c, const hpgb_geom &g, int64_t pix, double &z, double &phi, double &tcap,
                            bool &polar) {
    uint32_t nr, tphi, ring;  // ring size / 4, and the integer the reference converts for phi
    int32_t zb;               // belt: 2 nside - ring
    bool north, cap;
    double phi_off;
    const bool ok = hpgb_pix2zone_k<NEST>(c, pix, nr, tphi, zb, north, cap, phi_off, ring);
    const double dnr = hpgb_u2d(nr);
    tcap = (dnr * dnr) * g.fact2;  // RN(nr^2) either way: same double as (double)(nr * nr)
    const double zn = 1.0 - tcap;
    const double zc = north ? zn : -zn;  // tcap - 1 = -(1 - tcap) exactly
    z = cap ? zc : hpgb_i2d(zb) * g.fact1;
    polar = cap && nr <= c.nr_polar;  // <=> |zc| > 0.99
    phi = hpgb_zone_phi_k<NEST>(g, cap, tphi, phi_off, dnr);
    return ok;
}

// pixel -> (ring number, phi) only, for the kernels that take the ring-dependent outputs (colatitude, z, sin theta)
// from a per-launch ring table: same integers and the same phi as hpgb_pix2loc_k
template <bool NEST>
HPGB_HD bool hpgb_pix2ringphi_k(const hpgb_rk &c, const hpgb_geom &g, int64_t pix, uint32_t &ring, bool &north,
                                double &phi) {
    uint32_t nr, tphi;
    int32_t zb;
    bool cap;
    double phi_off;
    const bool ok = hpgb_pix2zone_k<NEST>(c, pix, nr, tphi, zb, north, cap, phi_off, ring);
    phi = hpgb_zone_phi_k<NEST>(g, cap, tphi, phi_off, hpgb_u2d(nr));
    return ok;
}

// pixel -> angles: hpgeom/healpix_geom.c:172-181 + hpgeom/hpgeom_utils.c:145-165.
// TA / TQ: the acos / atan tables of hpgb_math.h (table-driven, correctly rounded in practice).
template <bool NEST, bool POW2, bool LONLAT, bool DEGREES>
HPGB_HD void hpgb_pix2ang(const hpgb_geom &g, const uint64_t *TA, const double *TQ, int64_t pix, double &a, double &b) {
    double z, phi, sth;
    bool have_sth;
    hpgb_pix2loc<NEST, POW2>(g, pix, z, phi, sth, have_sth);
    // have_sth <=> |z| > 0.99 (hpgeom/healpix_geom.c:317,347-360): the reference takes atan2(sth, z) there
    const double theta = have_sth ? hpgb_atan2_tab(TQ, sth, z) : hpgb_acos_tab(TA, z);
    if (LONLAT) {
        double lon = phi;
        double lat = -(theta - HPGB_HALFPI);
        if (DEGREES) {
            lon *= 180.0 / HPGB_PI;
            lat *= 180.0 / HPGB_PI;
        }
        a = lon;
        b = lat;
    } else {
        a = theta;
        b = phi;
    }
}

// theta -> output pair (hpgeom/hpgeom_utils.c:145-165)
template <bool LONLAT, bool DEGREES>
HPGB_HD void hpgb_thetaphi_out(double theta, double phi, double &a, double &b) {
    if (LONLAT) {
        double lon = phi;
        double lat = -(theta - HPGB_HALFPI);
        if (DEGREES) {
            lon *= 180.0 / HPGB_PI;
            lat *= 180.0 / HPGB_PI;
        }
        a = lon;
        b = lat;
    } else {
        a = theta;
        b = phi;
    }
}

// pixel_to_angle for nside = 2^k, hot part: everything except the two rare cases, which the caller
// resolves afterwards (so that its pixels interleave): returns 0 = done, 1 = |z| > 0.99 (theta =
// atan2(sth, z), hpgb_pix2ang_polar), 2 = isqrt quirk (hpgb_pix2ang).
template <bool NEST, bool LONLAT, bool DEGREES>
HPGB_HD int hpgb_pix2ang_k(const hpgb_rk &c, const hpgb_geom &g, const uint64_t *TA, int64_t pix, double &a, double &b,
                           double &z, double &phi, double &tcap) {
    bool polar;
    const bool ok = hpgb_pix2loc_k<NEST>(c, g, pix, z, phi, tcap, polar);
    hpgb_thetaphi_out<LONLAT, DEGREES>(hpgb_acos_tab(TA, z), phi, a, b);
    return ok ? (polar ? 1 : 0) : 2;
}
template <bool LONLAT, bool DEGREES>
HPGB_HD void hpgb_pix2ang_polar(const double *TQ, double z, double phi, double tcap, double &a, double &b) {
    const double sth = sqrt(tcap * (2. - tcap));  // hpgeom/healpix_geom.c:317
    hpgb_thetaphi_out<LONLAT, DEGREES>(hpgb_atan2_tab(TQ, sth, z), phi, a, b);
}

// pixel -> unit vector: hpgeom/healpix_geom.c:183-199
template <bool NEST, bool POW2>
HPGB_HD void hpgb_pix2vec(const hpgb_geom &g, const double *T, int64_t pix, double &x, double &y, double &zz) {
    double z, phi, sth;
    bool have_sth;
    hpgb_pix2loc<NEST, POW2>(g, pix, z, phi, sth, have_sth);
    if (!have_sth) sth = sqrt((1 - z) * (1 + z));
    double s, c;
    hpgb_sincos_cr(T, phi, s, c);
    x = sth * c;
    y = sth * s;
    zz = z;
}

// pixel_to_vector for nside = 2^k, straight-line (see hpgb_pix2loc_k); false = isqrt quirk, use hpgb_pix2vec
template <bool NEST>
HPGB_HD bool hpgb_pix2vec_k(const hpgb_rk &c, const hpgb_geom &g, const double *T, int64_t pix, double &x, double &y,
                            double &zz) {
    double z, phi, tcap;
    bool polar;
    const bool ok = hpgb_pix2loc_k<NEST>(c, g, pix, z, phi, tcap, polar);
    // sin(theta): hpgeom/healpix_geom.c:317 (polar, from tmp) or :193 (from z); the argument is never 0
    // for a pixel centre, so the branch-free square root applies
    const double sth = hpgb_sqrt_tame(polar ? tcap * (2. - tcap) : (1 - z) * (1 + z));
    // cos(phi), sin(phi) in plain doubles (hpgb_sincos_p, 28 fp64 operations, <= 0.8 ulp): x, y end up within 2 ulp of
    // the reference's (which multiplies glibc's values by the same sth) -- measured over 1.6e7 pixels at nside 2^10 ..
    // 2^29: max 2 ulp, 99.5 % bit-identical.  A/B switches: HPGB_P2V_SINCOS_SD (<= 0.51 ulp, 40 operations: 99.87 %
    // identical, -7 % throughput), HPGB_P2V_DD (double-double, correctly rounded).
    double s, co;
#if defined(HPGB_P2V_DD)
    hpgb_sincos_cr(T, phi, s, co);
#elif defined(HPGB_P2V_SINCOS_SD)
    hpgb_sincos_sd(T, phi, s, co);
#else
    hpgb_sincos_p(T, phi, s, co);
#endif
    x = sth * co;
    y = sth * s;
    zz = z;
    return ok;
}

// ---------------------------------------------------------------------------------------------
// Ring tables (nside = 2^k <= 2^20, large batches).  Everything pixel_to_angle / pixel_to_vector derive from z --
// the colatitude (acos / atan2) resp. sin(theta) (sqrt) -- depends on the RING NUMBER only, and a map has 4 nside - 1
// rings against 12 nside^2 pixels.  A launch that converts many more pixels than there are rings first fills a
// table with one entry per ring, computed by the very routine the per-pixel path uses on the first pixel of the
// ring (so the values are the same doubles), and the streaming kernel gathers from it (the table -- <= 32 MB --
// stays in L2) instead of redoing ~40 fp64 operations per pixel.
// ---------------------------------------------------------------------------------------------
// first pixel (RING scheme) of ring `ring`, 1 <= ring <= 4 nside - 1   (hpgeom/healpix_geom.c:304-345 inverted)
HPGB_HD int64_t hpgb_ring_first_pixel(const hpgb_geom &g, int64_t ring) {
    const int64_t ns = g.nside;
    if (ring < ns) return 2 * ring * (ring - 1);
    if (ring <= 3 * ns) return g.ncap + (ring - ns) * 4 * ns;
    const int64_t nr = 4 * ns - ring;
    return g.npix - 2 * nr * (nr + 1);
}
// pixel_to_angle table entry: the theta-derived output (theta, or latitude in radians / degrees) of every pixel of
// the ring.  `gr` = geometry of the RING scheme for this nside.
template <bool LONLAT, bool DEGREES>
HPGB_HD double hpgb_ringtab_angle(const hpgb_geom &gr, const uint64_t *TA, const double *TQ, int64_t ring) {
    double va, vb;
    hpgb_pix2ang<false, true, LONLAT, DEGREES>(gr, TA, TQ, hpgb_ring_first_pixel(gr, ring), va, vb);
    return LONLAT ? vb : va;
}
// pixel_to_vector table entry for ring m <= 2 nside (northern half; z(4 nside - m) = -z(m) exactly and sin(theta)
// is the same double: hpgeom/healpix_geom.c:193,317,336)
HPGB_HD void hpgb_ringtab_zsth(const hpgb_geom &gr, int64_t ring, double &z, double &sth) {
    double phi;
    bool have_sth;
    hpgb_pix2loc<false, true>(gr, hpgb_ring_first_pixel(gr, ring), z, phi, sth, have_sth);
    if (!have_sth) sth = sqrt((1 - z) * (1 + z));
}
// pixel_to_angle / pixel_to_vector of one pixel given its table entries
template <bool NEST, bool LONLAT, bool DEGREES>
HPGB_HD void hpgb_pix2ang_tab(const hpgb_rk &c, const hpgb_geom &g, int64_t pix, uint32_t &ring, double &phi_out) {
    bool north;
    double phi;
    hpgb_pix2ringphi_k<NEST>(c, g, pix, ring, north, phi);
    phi_out = (LONLAT && DEGREES) ? phi * (180.0 / HPGB_PI) : phi;
}

// vector -> pixel: hpgeom/healpix_geom.c:161-170 (+ vec3_length, hpgeom/hpgeom_stack.c:396).
// No validation, as in the reference.  phi = atan2(y, x) is the only libm call.
template <bool NEST>
HPGB_HD int64_t hpgb_vec2pix(const hpgb_geom &g, const double *T, double x, double y, double z) {
    const double len = sqrt(x * x + y * y + z * z);
    const double xl = 1. / len;
    const double phi = ((x == 0.) && (y == 0.)) ? 0.0 : hpgb_atan2_any(T, y, x);
    const double nz = z * xl;
    if (fabs(nz) > 0.99) return hpgb_loc2pix_exact<NEST>(g, nz, phi, sqrt(x * x + y * y) * xl, true);
    return hpgb_loc2pix_exact<NEST>(g, nz, phi, 0, false);
}

// ---------------------------------------------------------------------------------------------
// neighbours: hpgeom/healpix_geom.c:890-979.  Order SW, W, NW, N, NE, E, SE, S; -1 = none.
// The reference's face / swap tables are packed into 64-bit words (4 bits per entry) so the
// lookup is a shift, not a memory access.
// ---------------------------------------------------------------------------------------------
template <bool NEST, bool POW2>
HPGB_HD int64_t hpgb_xyf2pix(const hpgb_geom &g, int32_t ix, int32_t iy, uint32_t face) {
    return NEST ? hpgb_xyf2nest(g.order, (uint32_t)ix, (uint32_t)iy, face) : hpgb_xyf2ring(g, ix, iy, face);
}

// nb_facearray[nbnum][face] (hpgeom/healpix_geom.c:892-900), 0xF = -1, nibble `face` of word `nbnum`
HPGB_HD uint64_t hpgb_nb_face_word(int nb) {
    switch (nb) {
        case 0: return 0x98BAFFFFBA98ull;   // S : 8,9,10,11,-1,-1,-1,-1,10,11,8,9
        case 1: return 0x8BA9BA984765ull;   // SE: 5,6,7,4,8,9,10,11,9,10,11,8
        case 2: return 0xFFFF4765FFFFull;   // E : -1 x4, 5,6,7,4, -1 x4
        case 3: return 0xA98BA98B7654ull;   // SW: 4,5,6,7,11,8,9,10,11,8,9,10
        case 4: return 0xBA9876543210ull;   // centre
        case 5: return 0x476532100321ull;   // NE: 1,2,3,0,0,1,2,3,5,6,7,4
        case 6: return 0xFFFF6547FFFFull;   // W : -1 x4, 7,4,5,6, -1 x4
        case 7: return 0x765421032103ull;   // NW: 3,0,1,2,3,0,1,2,4,5,6,7
        default: return 0x3210FFFF1032ull;  // N : 2,3,0,1,-1,-1,-1,-1,0,1,2,3
    }
}
// nb_swaparray[nbnum][face>>2] (hpgeom/healpix_geom.c:901-909), 3 nibbles per row
HPGB_HD uint32_t hpgb_nb_swap(int nb, uint32_t frow) {
    // rows: S{0,0,3} SE{0,0,6} E{0,0,0} SW{0,0,5} C{0,0,0} NE{5,0,0} W{0,0,0} NW{6,0,0} N{3,0,0}
    const uint32_t tab[9] = {0x300u, 0x600u, 0x000u, 0x500u, 0x000u, 0x005u, 0x000u, 0x006u, 0x003u};
    return (tab[nb] >> (4u * frow)) & 7u;
}

// NEST pixel strictly inside its base face (0 < ix, iy < nside-1): the eight neighbours are
// obtained by +-1 on the even (ix) and odd (iy) bit planes of the Morton code itself -- dilated-
// integer arithmetic, no de-interleave / re-interleave.  Returns false (out untouched) for a pixel
// on a face edge; those (4/nside of all pixels) take the general path below.
HPGB_HD bool hpgb_neighbors_nest_interior(const hpgb_geom &g, int64_t pix, int64_t *out) {
    const uint64_t m = (uint64_t)(g.npface - 1);
    const uint64_t EV = 0x5555555555555555ull & m, OD = 0xAAAAAAAAAAAAAAAAull & m;
    const uint64_t v = (uint64_t)pix & m, fb = (uint64_t)pix - v;
    const uint64_t X = v & EV, Y = v & OD;
    if (X == 0 || X == EV || Y == 0 || Y == OD) return false;
    const uint64_t xp = ((X | OD) + 1) & EV;  // ix + 1: the carry rides over the odd plane
    const uint64_t xm = (X - 1) & EV;         // ix - 1: the borrow rides over the (zero) odd plane
    const uint64_t yp = ((Y | EV) + 2) & OD;
    const uint64_t ym = (Y - 2) & OD;
    out[0] = (int64_t)(fb + (xm | Y));   // SW (-1,  0)
    out[1] = (int64_t)(fb + (xm | yp));  // W  (-1, +1)
    out[2] = (int64_t)(fb + (X | yp));   // NW ( 0, +1)
    out[3] = (int64_t)(fb + (xp | yp));  // N  (+1, +1)
    out[4] = (int64_t)(fb + (xp | Y));   // NE (+1,  0)
    out[5] = (int64_t)(fb + (xp | ym));  // E  (+1, -1)
    out[6] = (int64_t)(fb + (X | ym));   // SE ( 0, -1)
    out[7] = (int64_t)(fb + (xm | ym));  // S  (-1, -1)
    return true;
}

template <bool NEST, bool POW2>
HPGB_HD void hpgb_neighbors(const hpgb_geom &g, int64_t pix, int64_t *out) {
    int32_t ix, iy;
    uint32_t face;
    if (NEST && hpgb_neighbors_nest_interior(g, pix, out)) return;
    if (NEST) {
        uint32_t ux, uy;
        hpgb_nest2xyf(g, pix, ux, uy, face);
        ix = (int32_t)ux;
        iy = (int32_t)uy;
    } else {
        hpgb_ring2xyf<POW2>(g, pix, ix, iy, face);
    }
    const int32_t ns = (int32_t)g.nside;
    const int32_t dx[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
    const int32_t dy[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    if (ix > 0 && ix < ns - 1 && iy > 0 && iy < ns - 1) {
#pragma unroll
        for (int m = 0; m < 8; m++) out[m] = hpgb_xyf2pix<NEST, POW2>(g, ix + dx[m], iy + dy[m], face);
        return;
    }
#pragma unroll
    for (int m = 0; m < 8; m++) {
        int32_t x = ix + dx[m], y = iy + dy[m];
        int nb = 4;
        if (x < 0) {
            x += ns;
            nb--;
        } else if (x >= ns) {
            x -= ns;
            nb++;
        }
        if (y < 0) {
            y += ns;
            nb -= 3;
        } else if (y >= ns) {
            y -= ns;
            nb += 3;
        }
        const uint32_t f = (uint32_t)((hpgb_nb_face_word(nb) >> (4u * face)) & 0xFull);
        if (f == 0xFu) {
            out[m] = -1;
            continue;
        }
        const uint32_t bits = hpgb_nb_swap(nb, face >> 2);
        if (bits & 1u) x = ns - x - 1;
        if (bits & 2u) y = ns - y - 1;
        if (bits & 4u) {
            const int32_t t = x;
            x = y;
            y = t;
        }
        out[m] = hpgb_xyf2pix<NEST, POW2>(g, x, y, f);
    }
}

// RING neighbours for nside = 2^k (scalar-nside kernel).  A pixel strictly inside its base face has its
// eight neighbours on five consecutive rings (N: ring-2; NW, NE: ring-1; W, E: same ring; SW, SE: ring+1;
// S: ring+2), so the zone / multiplier work of xyf2ring (hpgb_xyf2ring_k, hpgb_reorder.h: pixel = A*B + C)
// is done once per RING and only the in-ring offset once per neighbour; everything is selects.  Face-edge
// pixels (4/nside of all) take the general routine out of line.  Same values as hpgb_neighbors<false, true>.
HPGB_HD_NOINLINE void hpgb_neighbors_ring_edge(const hpgb_geom &g, int64_t pix, int64_t *out) {
    hpgb_neighbors<false, true>(g, pix, out);
}
// NEST: Morton-plane arithmetic for face-interior pixels, the general routine out of line for the rest
HPGB_HD_NOINLINE void hpgb_neighbors_nest_edge(const hpgb_geom &g, int64_t pix, int64_t *out) {
    hpgb_neighbors<true, true>(g, pix, out);
}
HPGB_HD void hpgb_neighbors_nest_k(const hpgb_geom &g, int64_t pix, int64_t *out) {
    if (hpgb_neighbors_nest_interior(g, pix, out)) return;
    int64_t tmp[8];
    hpgb_neighbors_nest_edge(g, pix, tmp);
#pragma unroll
    for (int m = 0; m < 8; m++) out[m] = tmp[m];
}
HPGB_HD void hpgb_neighbors_ring_k(const hpgb_rk &c, const hpgb_geom &g, int64_t pix, int64_t *out) {
    uint32_t ix, iy, face;
    hpgb_ring2xyf_k(c, g, pix, ix, iy, face);
    if (!(ix > 0u && ix < c.nsm1 && iy > 0u && iy < c.nsm1)) {
        int64_t tmp[8];  // the out-of-line call gets its own (local-memory) row, so `out` stays in registers
        hpgb_neighbors_ring_edge(g, pix, tmp);
#pragma unroll
        for (int m = 0; m < 8; m++) out[m] = tmp[m];
        return;
    }
    const uint32_t fc = face & 3u, fr = face >> 2;
    const uint32_t frns = fr * c.ns;
    const uint32_t t0 = frns + c.nsm1 - ix - iy;                            // ring - nside of the centre
    const uint32_t w0 = fc * c.ns2 + (ix - iy) + c.nsm1 - (frns & c.ns);    // jpll * nside + ix - iy - 1
    const uint32_t fm2 = fc - 2u;
    uint32_t A[5], B[5], Ch[5], zn[5], zs[5], ks[5];
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const uint32_t t = t0 + (uint32_t)(r - 2);
        zn[r] = (int32_t)t < 0 ? 1u : 0u;
        zs[r] = (int32_t)t >= (int32_t)c.ns2 ? 1u : 0u;
        const uint32_t nr = hpgb_sel(zn[r], t + c.ns, c.ns3 - t);
        A[r] = hpgb_sel(zn[r] | zs[r], nr, t + c.belt_a0);
        B[r] = hpgb_sel(zn[r], 2u * nr + fm2, hpgb_sel(zs[r], fm2 - 2u * nr, c.ns4));
        Ch[r] = hpgb_sel(zs[r], c.npix_hi - nr, 0u);
        ks[r] = t & 1u;
    }
    // neighbour order SW, W, NW, N, NE, E, SE, S (hpgeom/hpgeom.c:2879); ring index = 2 - (dx + dy)
    const int dx[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
    const int dy[8] = {0, 1, 1, 1, 0, -1, -1, -1};
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int r = 2 - (dx[m] + dy[m]);
        const uint32_t X = ((w0 + (uint32_t)(dx[m] - dy[m]) + ks[r]) >> 1) & c.ns4m1;
        const uint32_t cl = hpgb_sel(zn[r], ~(iy + (uint32_t)dy[m]) & c.nsm1,
                                     hpgb_sel(zs[r], c.npix_lo | (ix + (uint32_t)dx[m]), X + c.belt_c0));
        out[m] = (int64_t)((uint64_t)A[r] * (uint64_t)B[r] + (((uint64_t)Ch[r] << 32) | cl));
    }
}

// ---------------------------------------------------------------------------------------------
// bilinear interpolation: hpgeom/healpix_geom.c:504-510 (ring_above), :1585-1606
// (get_ring_info2), :1608-1690 (get_interpol)
// ---------------------------------------------------------------------------------------------
HPGB_HD int64_t hpgb_ring_above(const hpgb_geom &g, double z) {
    const double az = fabs(z);
    const double dns = (double)g.nside;
    if (az <= 2.0 / 3.0) return hpgb_d2ll(dns * (2 - 1.5 * z));
    const int64_t ir = hpgb_d2ll(dns * sqrt(3 * (1 - az)));
    return (z > 0) ? ir : 4 * g.nside - ir - 1;
}

// The arithmetic the interpolation needs beyond + - *: two interchangeable providers.
//   hpgb_mp_ref : IEEE / and sqrt, double-double acos / atan2 (any nside; the reference-shaped path)
//   hpgb_mp_tab : the same values from the table-driven acos / atan2 of hpgb_math.h and the
//                 branch-free division / square root (power-of-two nside kernels)
struct hpgb_mp_ref {
    const double *T;
    HPGB_HDM double acos_(double z) const { return hpgb_acos_cr(T, z); }
    HPGB_HDM double atan2_(double s, double c) const { return hpgb_atan2_cr(T, s, c); }
    HPGB_HDM double div_(double a, double b) const { return a / b; }
    HPGB_HDM double sqrt_(double x) const { return sqrt(x); }
};
struct hpgb_mp_tab {
    const uint64_t *TA;
    const double *TQ;
    HPGB_HDM double acos_(double z) const { return hpgb_acos_tab(TA, z); }
    HPGB_HDM double atan2_(double s, double c) const { return hpgb_atan2_tab(TQ, s, c, HPGB_ATANQ_ROWS); }
    HPGB_HDM double div_(double a, double b) const { return hpgb_div_tame(a, b); }
    HPGB_HDM double sqrt_(double x) const { return hpgb_sqrt_tame(x); }
};

// Ring numbers and ring lengths fit 32 bits for every legal nside (ring < 4 nside <= 2^31, ring length
// <= 2^31), so the integer work is 32-bit with 32x32->64 products; (double)nring * (double)nring is the
// correctly rounded nring^2, i.e. the reference's (double)(nring * nring).
template <class MP>
HPGB_HD void hpgb_ring_info2(const hpgb_geom &g, const MP &mp, int64_t ring, int64_t &startpix, int64_t &ringpix,
                             double &theta, bool &shifted) {
    const uint32_t ns = (uint32_t)g.nside, r = (uint32_t)ring;
    const uint32_t nring = (r > 2u * ns) ? 4u * ns - r : r;
    uint32_t rp;
    if (nring < ns) {
        const double dn = (double)nring;
        const double t = (dn * dn) * g.fact2;
        const double ct = 1 - t;
        const double st = mp.sqrt_(t * (2 - t));
        theta = mp.atan2_(st, ct);
        rp = 4u * nring;
        shifted = true;
        startpix = (int64_t)((uint64_t)(2u * nring) * (uint64_t)(nring - 1u));
    } else {
        theta = mp.acos_((double)(2u * ns - nring) * g.fact1);
        rp = 4u * ns;
        shifted = ((nring - ns) & 1u) == 0u;
        startpix = g.ncap + (int64_t)((uint64_t)(nring - ns) * (uint64_t)rp);
    }
    ringpix = (int64_t)rp;
    if (nring != r) {
        theta = HPGB_PI - theta;
        startpix = g.npix - startpix - ringpix;
    }
}

// one ring: two pixels + phi weights (the twin blocks hpgeom/healpix_geom.c:1619-1636, :1637-1654).
// NESTK: write NEST pixel numbers straight from (ring, in-ring index) -- power-of-two nside only.
template <bool NESTK, class MP>
HPGB_HD void hpgb_ring_pair(const hpgb_geom &g, const hpgb_rk &rk, const MP &mp, int64_t ring, double phi, int64_t *p,
                            double *w, double &theta) {
    int64_t sp, nr;
    bool shift;
    hpgb_ring_info2(g, mp, ring, sp, nr, theta, shift);
    const double dphi = mp.div_(HPGB_TWOPI, (double)(uint32_t)nr);
    const double hs = shift ? 0.5 : 0.0;
    const double t = (mp.div_(phi, dphi) - hs);
    int64_t i1 = (t < 0) ? hpgb_d2ll(t) - 1 : hpgb_d2ll(t);
    const double w1 = mp.div_(phi - ((double)i1 + hs) * dphi, dphi);
    int64_t i2 = i1 + 1;
    if (i1 < 0) i1 += nr;
    if (i2 >= nr) i2 -= nr;
    if (NESTK) {
        hpgb_ringidx2nest_pair_k(rk, (uint32_t)ring, (uint32_t)i1, (uint32_t)i2, p[0], p[1]);
    } else {
        p[0] = sp + i1;
        p[1] = sp + i2;
    }
    w[0] = 1 - w1;
    w[1] = w1;
}

// everything after the choice of the ring pair (hpgeom/healpix_geom.c:1612-1690); ir1 = ring above the point.
// rk: constants of the power-of-two RING -> NEST core (POW2 only)
template <bool NEST, bool POW2, class MP>
HPGB_HD void hpgb_interpol_rings(const hpgb_geom &g, const hpgb_rk &rk, const MP &mp, double theta, double phi,
                                 int64_t ir1, int64_t *p, double *w) {
    const int64_t ir2 = ir1 + 1;
    double th1 = 0, th2 = 0;
    p[0] = p[1] = p[2] = p[3] = 0;
    w[0] = w[1] = w[2] = w[3] = 0;
    if (NEST && POW2 && ir1 > 0 && ir2 < 4 * g.nside) {  // the common case: NEST numbers straight from (ring, index)
        hpgb_ring_pair<true>(g, rk, mp, ir1, phi, p, w, th1);
        hpgb_ring_pair<true>(g, rk, mp, ir2, phi, p + 2, w + 2, th2);
        const double wt = mp.div_(theta - th1, th2 - th1);
        w[0] *= (1 - wt);
        w[1] *= (1 - wt);
        w[2] *= wt;
        w[3] *= wt;
        return;
    }
    if (ir1 > 0) hpgb_ring_pair<false>(g, rk, mp, ir1, phi, p, w, th1);
    if (ir2 < 4 * g.nside) hpgb_ring_pair<false>(g, rk, mp, ir2, phi, p + 2, w + 2, th2);
    if (ir1 == 0) {
        const double wt = mp.div_(theta, th2);
        w[2] *= wt;
        w[3] *= wt;
        const double fac = (1 - wt) * 0.25;
        w[0] = fac;
        w[1] = fac;
        w[2] += fac;
        w[3] += fac;
        p[0] = (p[2] + 2) & 3;
        p[1] = (p[3] + 2) & 3;
    } else if (ir2 == 4 * g.nside) {
        const double wt = mp.div_(theta - th1, HPGB_PI - th1);
        w[0] *= (1 - wt);
        w[1] *= (1 - wt);
        const double fac = wt * 0.25;
        w[0] += fac;
        w[1] += fac;
        w[2] = fac;
        w[3] = fac;
        p[2] = ((p[0] + 2) & 3) + g.npix - 4;
        p[3] = ((p[1] + 2) & 3) + g.npix - 4;
    } else {
        const double wt = mp.div_(theta - th1, th2 - th1);
        w[0] *= (1 - wt);
        w[1] *= (1 - wt);
        w[2] *= wt;
        w[3] *= wt;
    }
    if (NEST) {
#pragma unroll
        for (int m = 0; m < 4; m++) p[m] = POW2 ? hpgb_ring2nest_k<false>(rk, g, p[m]) : hpgb_ring2nest(g, p[m]);
    }
}

// reference-shaped: z = cos(theta) (correctly rounded), ring_above(z)
template <bool NEST, bool POW2>
HPGB_HD void hpgb_interpol(const hpgb_geom &g, const double *T, double theta, double phi, int64_t *p, double *w) {
    const double z = hpgb_cos_cr(T, theta);
    const hpgb_mp_ref mp{T};
    hpgb_interpol_rings<NEST, false>(g, hpgb_rk(), mp, theta, phi, hpgb_ring_above(g, z), p, w);
}

// Fast choice of the ring pair for nside = 2^k: ring_above(cos theta) is a step function of theta, so
// away from its steps a cheap cosine decides it (the polynomial sine of the angle_to_pixel fast path:
// z = sin(pi/2 - theta) in the belt, nside sqrt(3 (1 - |z|)) = nside sqrt(6) sin(colat/2) in the caps).
// Returns false -- the caller then runs hpgb_interpol -- when the value is within nside * 2^-44 of a
// step, near the zone transition, or within 0.0102 rad of a pole (where the reference's own cos /
// 1 - |z| is noisy, and where the two pole special cases live).
HPGB_HD bool hpgb_ring_above_fast(const hpgb_geom &g, const hpgb_a2p_fast &f, double theta, int64_t &ir1) {
    const double lat = (HPGB_PIO2_1 - theta) + HPGB_PIO2_2;
    const bool north = theta < HPGB_PIO2_1;
    const double colat = north ? theta : (HPGB_PI_1 - theta) + HPGB_PI_2;  // from the nearest pole
    const double al = fabs(lat);
    const bool eq = al <= f.asin23;
    bool ok = (colat > 0.0102) & !(fabs(al - f.asin23) < 1e-11);
    const double s = hpgb_sin_poly(f, eq ? lat : 0.5 * colat);
    const double dns = (double)g.nside;
    const double v = eq ? dns * (2.0 - 1.5 * s) : dns * (2.4494897427831780981972840747059 * s);
    const double fl = floor(v);
    const double fr = v - fl, m = dns * 5.6843418860808015e-14;  // 2^-44
    ok &= (fr > m) & (fr < 1.0 - m);
    const int64_t iv = (int64_t)fl;
    ir1 = (eq | north) ? iv : 4 * g.nside - iv - 1;
    return ok;
}

// ---------------------------------------------------------------------------------------------
// Interpolation, zone by zone (the sorted kernel of k_interp.cu).  hpgb_interpol_rings serves any ring pair
// with per-ring branches: a warp whose 32 points straddle the caps and the belt runs the acos AND the atan2
// colatitude, and both ring-length formulas, for every point.  The kernel therefore sorts the points of a warp
// by the zone of their ring pair and calls one of the two straight-line bodies below on 32 points of ONE zone.
// Same arithmetic, operation for operation, as hpgb_ring_pair / hpgb_ring_info2 with the hpgb_mp_tab
// providers (hpgeom/healpix_geom.c:1585-1690); only what a zone makes constant is hoisted:
//   belt pair (nside <= ir1, ir1 + 1 <= 3 nside): both rings have 4 nside pixels, so dphi = 2 pi / (4 nside) is
//     a per-nside constant (computed on the host with the same IEEE division) and phi / dphi is shared;
//   cap pair (both rings strictly inside one polar cap): ring colatitude from atan2, ring length 4 * ring.
// Everything else (the two pole rows, a pair that straddles a cap edge, a point too close to a ring for the
// cheap ring choice) is rare and goes to hpgb_interpol / hpgb_interpol_rings.
// ---------------------------------------------------------------------------------------------
// the branch-free divisions of the zone bodies (hpgb_div_tame, hpgb_div_const) are IEEE-exact away from underflow;
// phi / dphi with a subnormal-range phi is not: such points (0 < phi < 1e-290) go to the reference-shaped routine
HPGB_HD bool hpgb_interp_phi_tame(double phi) { return !(phi > 0.0 && phi < 1e-290); }
// zone of the pair (ir1, ir1 + 1): 0 belt, 1 cap, 2 anything else
HPGB_HD int hpgb_interp_zone(uint32_t ns, int64_t ir1) {
    const uint32_t r = (uint32_t)ir1;
    if (ir1 >= (int64_t)ns && ir1 + 1 <= 3 * (int64_t)ns) return 0;
    if ((ir1 >= 1 && ir1 + 1 <= (int64_t)ns - 1) || (ir1 >= 3 * (int64_t)ns + 1 && ir1 + 1 <= 4 * (int64_t)ns - 1)) return 1;
    (void)r;
    return 2;
}

// hpgb_ring_above_fast + hpgb_interp_zone for the sorted kernel, in 32-bit integers and without conversion-unit
// instructions: floor(v) is RN(v - 1/2) read off the mantissa of (v - 1/2) + 1.5 2^52 (v - 1/2 is exact for v >= 1 and
// the two disagree only where frac(v) is 0 -- inside the guard band, which then fails), `dns` = (double)nside is
// computed once per kernel, the zone is three unsigned compares.  Returns 0 = both rings in the belt, 1 = both in
// one polar cap, 2 = anything else (guard failed, pole rows, pairs straddling a cap edge): same ring and the same
// verdict as the two routines above (tests/test_hostsim.py runs both).
HPGB_HD int hpgb_ring_zone_fast(const hpgb_rk &rk, const hpgb_a2p_fast &f, double dns, double theta, uint32_t &ir1) {
    const double lat = (HPGB_PIO2_1 - theta) + HPGB_PIO2_2;
    const bool north = theta < HPGB_PIO2_1;
    const double colat = north ? theta : (HPGB_PI_1 - theta) + HPGB_PI_2;  // from the nearest pole
    const double al = fabs(lat);
    const bool eq = al <= f.asin23;
    bool ok = (colat > 0.0102) & !(fabs(al - f.asin23) < 1e-11);
    const double s = hpgb_sin_poly(f, eq ? lat : 0.5 * colat);
    const double v = eq ? dns * (2.0 - 1.5 * s) : dns * (2.4494897427831780981972840747059 * s);
    const double t = (v - 0.5) + 6755399441055744.0;
    const double fl = t - 6755399441055744.0;
    const double fr = v - fl, m = dns * 5.6843418860808015e-14;  // 2^-44
    ok &= (fr > m) & (fr < 1.0 - m);
    const uint32_t iv = hpgb_lo32(t);
    ir1 = (eq | north) ? iv : rk.ns4 - iv - 1u;
    const bool belt = (ir1 - rk.ns) <= rk.ns2 - 1u;
    const uint32_t capw = rk.ns > 2u ? rk.ns - 2u : 0u;
    const bool cap = (ir1 - 1u) < capw || (ir1 - rk.ns3p1) < capw;
    return ok ? (belt ? 0 : (cap ? 1 : 2)) : 2;
}

// floor of a value >= -1 as the reference forms it: (t < 0) ? (int64)t - 1 : (int64)t  (hpgeom/healpix_geom.c:1622)
// (a ring of nside 2^29 has 2^31 pixels: the index needs 64 bits, like the reference's)
HPGB_HD int64_t hpgb_interp_floor(double t) { return (t < 0) ? hpgb_d2ll(t) - 1 : hpgb_d2ll(t); }

template <bool NEST>
HPGB_HD void hpgb_interp_belt(const hpgb_geom &g, const hpgb_rk &rk, const uint64_t *TA, double dphi, double theta, double phi,
                              uint32_t ir1, int64_t *p, double *w) {
    const uint32_t ns = (uint32_t)g.nside, nr = 4u * ns;
    const double tc = hpgb_div_tame(phi, dphi);  // shared by the two rings
    double th[2], w1[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const uint32_t r = ir1 + (uint32_t)k;
        const bool south = r > 2u * ns;
        const uint32_t nring = south ? 4u * ns - r : r;
        const double t0 = hpgb_acos_tab(TA, (double)(2u * ns - nring) * g.fact1);
        th[k] = south ? HPGB_PI - t0 : t0;
        const double hs = ((nring - ns) & 1u) == 0u ? 0.5 : 0.0;
        int64_t i1 = hpgb_interp_floor(tc - hs);
        w1[k] = hpgb_div_tame(phi - ((double)i1 + hs) * dphi, dphi);
        int64_t i2 = i1 + 1;
        if (i1 < 0) i1 += (int64_t)nr;
        if (i2 >= (int64_t)nr) i2 -= (int64_t)nr;
        if (NEST) {
            hpgb_ringidx2nest_pair_k(rk, r, (uint32_t)i1, (uint32_t)i2, p[2 * k], p[2 * k + 1]);
        } else {
            int64_t sp = g.ncap + (int64_t)((uint64_t)(nring - ns) * (uint64_t)nr);
            if (south) sp = g.npix - sp - (int64_t)nr;
            p[2 * k] = sp + i1;
            p[2 * k + 1] = sp + i2;
        }
    }
    const double wt = hpgb_div_tame(theta - th[0], th[1] - th[0]);
    w[0] = (1 - w1[0]) * (1 - wt);
    w[1] = w1[0] * (1 - wt);
    w[2] = (1 - w1[1]) * wt;
    w[3] = w1[1] * wt;
}

template <bool NEST>
HPGB_HD void hpgb_interp_cap(const hpgb_geom &g, const hpgb_rk &rk, const double *TQ, double theta, double phi, uint32_t ir1,
                             int64_t *p, double *w) {
    const uint32_t ns = (uint32_t)g.nside;
    double th[2], w1[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const uint32_t r = ir1 + (uint32_t)k;
        const bool south = r > 2u * ns;
        const uint32_t nring = south ? 4u * ns - r : r;
        const double dn = (double)nring;
        const double t = (dn * dn) * g.fact2;
        const double t0 = hpgb_atan2_tab(TQ, hpgb_sqrt_tame(t * (2 - t)), 1 - t, HPGB_ATANQ_ROWS);
        th[k] = south ? HPGB_PI - t0 : t0;
        const uint32_t nr = 4u * nring;
        const double dphi = hpgb_div_tame(HPGB_TWOPI, (double)nr);
        int64_t i1 = hpgb_interp_floor(hpgb_div_tame(phi, dphi) - 0.5);
        w1[k] = hpgb_div_tame(phi - ((double)i1 + 0.5) * dphi, dphi);
        int64_t i2 = i1 + 1;
        if (i1 < 0) i1 += (int64_t)nr;
        if (i2 >= (int64_t)nr) i2 -= (int64_t)nr;
        if (NEST) {
            hpgb_ringidx2nest_pair_k(rk, r, (uint32_t)i1, (uint32_t)i2, p[2 * k], p[2 * k + 1]);
        } else {
            int64_t sp = (int64_t)((uint64_t)(2u * nring) * (uint64_t)(nring - 1u));
            if (south) sp = g.npix - sp - (int64_t)nr;
            p[2 * k] = sp + i1;
            p[2 * k + 1] = sp + i2;
        }
    }
    const double wt = hpgb_div_tame(theta - th[0], th[1] - th[0]);
    w[0] = (1 - w1[0]) * (1 - wt);
    w[1] = w1[0] * (1 - wt);
    w[2] = (1 - w1[1]) * wt;
    w[3] = w1[1] * wt;
}

// ---------------------------------------------------------------------------------------------
// Interpolation for SMALL nside (<= HPGB_INTERP_LUT_MAX_NSIDE), third generation: what depends only on the RING is
// computed once per launch into a shared-memory table instead of once per point --
//   th[nring]            colatitude of ring nring, 1 <= nring <= 2 nside (the southern rings are pi - th, as in the
//                        reference's get_ring_info2, hpgeom/healpix_geom.c:1585-1610)
//   cap[nring] = (d, inv) d = 2 pi / (4 nring), the pixel spacing of a polar-cap ring, and inv = RN(1 / d)
// -- by the SAME routines the per-point bodies above call (hpgb_acos_tab / hpgb_atan2_tab / IEEE division), so every
// value is bit-identical; the table is the reference's "ring info" hoisted out of the point loop (two correctly
// rounded inverse trig evaluations, ~200 instructions, leave the body of a point).
// Also here: 32-bit ring indices (a ring has at most 4 nside <= 2^14 pixels), NEST numbers through the folded
// coordinates of hpgb_reorder.h with a ONE-word Morton step (X, Y < 2^14), and divisions by the per-ring constant
// d as q = RN(a inv), q' = RN(q + (a - d q) inv) -- correctly rounded when inv = RN(1/d) and d's significand is not
// all ones (Markstein's theorem; d = pi / (2 nring)); checked against IEEE division in tests/test_hostsim.py.
// ---------------------------------------------------------------------------------------------
#define HPGB_INTERP_LUT_MAX_NSIDE 4096
struct hpgb_cdiv {
    double d, inv;
};
HPGB_HD double hpgb_div_const(double a, const hpgb_cdiv &c) {
    const double q = a * c.inv;
    return fma(fma(-c.d, q, a), c.inv, q);
}
HPGB_HD int32_t hpgb_d2i(double v) {
#if defined(__CUDA_ARCH__)
    return __double2int_rz(v);
#else
    return (int32_t)v;
#endif
}
// one table entry (run by the kernel's prologue and by the host simulation): ring nring in 1 .. 2 nside
HPGB_HD void hpgb_interp_ring_entry(const hpgb_geom &g, const uint64_t *TA, const double *TQ, uint32_t nring, double &theta,
                                    hpgb_cdiv &cap) {
    const uint32_t ns = (uint32_t)g.nside;
    cap.d = cap.inv = 0.0;
    if (nring < ns) {
        const double dn = (double)nring;
        const double t = (dn * dn) * g.fact2;
        theta = hpgb_atan2_tab(TQ, hpgb_sqrt_tame(t * (2 - t)), 1 - t, HPGB_ATANQ_ROWS);
        cap.d = HPGB_TWOPI / (double)(4u * nring);
        cap.inv = 1.0 / cap.d;
    } else {
        theta = hpgb_acos_tab(TA, (double)(2u * ns - nring) * g.fact1);
    }
}
// Morton code of two coordinates below 2^16 in one word
HPGB_HD uint32_t hpgb_interleave16(uint32_t x, uint32_t y) {
    uint32_t lo = hpgb_prmt(x, y, 0x5140);
    HPGB_DSWAP(lo, 0x00F000F0u, 4);
    HPGB_DSWAP(lo, 0x0C0C0C0Cu, 2);
    HPGB_DSWAP(lo, 0x22222222u, 1);
    return lo;
}
// (belt ring r, nside <= r <= 3 nside; 0-based in-ring index q) -> NEST pixel; flut = the 16 (lx, ly) pairs of
// hpgb_r2n_lut_entry
HPGB_HD int64_t hpgb_beltidx2nest_fold(const hpgb_rk &c, const hpgb_u32x2 *flut, uint32_t r, uint32_t q) {
    uint32_t X, Y;
    hpgb_r2n_belt_fold(c, flut, r - c.ns, q, X, Y);  // q < 4 nside: the mask inside is a no-op
    return (int64_t)hpgb_interleave16(X, Y);
}
// (cap ring: nring < nside from the nearest pole, south or north; 0-based in-ring index q) -> NEST pixel
HPGB_HD int64_t hpgb_capidx2nest_fold(const hpgb_rk &c, uint32_t nring, bool south, uint32_t q) {
    const bool t1 = q >= 2u * nring;
    q -= t1 ? 2u * nring : 0u;
    const bool t2 = q >= nring;
    q -= t2 ? nring : 0u;
    const uint32_t X = (south ? q : c.ns - nring + q) + (t2 ? c.ns : 0u);
    const uint32_t Y = (south ? nring - 1u - q : c.nsm1 - q) + (t1 ? c.ns : 0u) + (south ? c.ns2 : 0u);
    return (int64_t)hpgb_interleave16(X, Y);
}
// floor of t >= -1/2 the way the reference forms it, 32-bit
HPGB_HD int32_t hpgb_interp_floor32(double t) { return (t < 0) ? hpgb_d2i(t) - 1 : hpgb_d2i(t); }

template <bool NEST>
HPGB_HD void hpgb_interp_belt_lut(const hpgb_geom &g, const hpgb_rk &rk, const hpgb_u32x2 *flut, const double *th_tab,
                                  const hpgb_cdiv &dphi, double theta, double phi, uint32_t ir1, int64_t *p, double *w) {
    const uint32_t ns = rk.ns, nr = rk.ns4;
    const double tc = hpgb_div_const(phi, dphi);  // shared by the two rings
    double th[2], w1[2];
    int32_t ii[2] = {0, 0};
    uint32_t qq[4] = {0u, 0u, 0u, 0u};
    bool sw = false;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const uint32_t r = ir1 + (uint32_t)k;
        const bool south = r > rk.ns2;
        const uint32_t nring = south ? nr - r : r;
        const double t0 = th_tab[nring];
        th[k] = south ? HPGB_PI - t0 : t0;
        const double hs = ((nring - ns) & 1u) == 0u ? 0.5 : 0.0;
        const int32_t i1 = hpgb_interp_floor32(tc - hs);
        w1[k] = hpgb_div_const(phi - ((double)i1 + hs) * dphi.d, dphi);
        const uint32_t q1 = i1 < 0 ? (uint32_t)i1 + nr : (uint32_t)i1;
        const uint32_t q2 = (uint32_t)(i1 + 1) >= nr ? (uint32_t)(i1 + 1) - nr : (uint32_t)(i1 + 1);
        if (NEST) {
#if defined(HPGB_INTERP_NO_DIAMOND)
            p[2 * k] = hpgb_beltidx2nest_fold(rk, flut, r, q1);
            p[2 * k + 1] = hpgb_beltidx2nest_fold(rk, flut, r, q2);
#else
            ii[k] = i1;
            qq[2 * k] = q1;
            qq[2 * k + 1] = q2;
            sw = k == 0 ? (hs != 0.0) : sw;  // hs of ring r
#endif
        } else {
            uint32_t sp = rk.ncap_lo + (nring - ns) * nr;
            if (south) sp = rk.npix_lo - sp - nr;
            p[2 * k] = (int64_t)(sp + q1);
            p[2 * k + 1] = (int64_t)(sp + q2);
        }
    }
#if !defined(HPGB_INTERP_NO_DIAMOND)
    if (NEST) {
        // The four pixels are a diamond of the pixel grid: with (ix, iy) the in-face coordinates of the first one,
        // its ring mate is the EAST neighbour (ix + 1, iy - 1), the first pixel of the ring below is its SOUTH-WEST
        // (ix - 1, iy) or SOUTH-EAST (ix, iy - 1) neighbour -- one ring further south the edge-line coordinates
        // change by jm - 1 (ring - nside even) or jp + 1 (odd) at equal in-ring index, and the index itself moves by
        // 0 / +1 resp. -1 / 0 because the rings are shifted by half a pixel against each other -- and the fourth
        // is that one's east neighbour.  Inside a base face each step is +-1 on one bit plane of the Morton code
        // (the face bits ride above, hpgb_reorder.h "folded coordinates"): one conversion and three 5-instruction
        // steps instead of four conversions.  A step that would leave the face (2-3 pixels in nside) converts
        // directly.
        const uint32_t EV = 0x55555555u & rk.lmask, OD = 0xAAAAAAAAu & rk.lmask;
        const uint32_t v0 = (uint32_t)hpgb_beltidx2nest_fold(rk, flut, ir1, qq[0]);
        const bool use_sw = sw ? (ii[1] == ii[0]) : (ii[1] != ii[0]);
        // east of v: x + 1, y - 1 (needs x < nside - 1, y > 0)
        const uint32_t e0 = (v0 & ~rk.lmask) | (((v0 | ~EV) + 1u) & EV) | (((v0 & OD) - 2u) & OD);
        const bool e0_ok = (v0 & EV) != EV && (v0 & OD) != 0u;
        // south-west: x - 1 (needs x > 0); south-east: y - 1 (needs y > 0)
        const uint32_t s0 = use_sw ? ((v0 & ~EV) | (((v0 & EV) - 1u) & EV)) : ((v0 & ~OD) | (((v0 & OD) - 2u) & OD));
        const bool s0_ok = use_sw ? ((v0 & EV) != 0u) : ((v0 & OD) != 0u);
        const uint32_t e1 = (s0 & ~rk.lmask) | (((s0 | ~EV) + 1u) & EV) | (((s0 & OD) - 2u) & OD);
        const bool e1_ok = s0_ok && (s0 & EV) != EV && (s0 & OD) != 0u;
        p[0] = (int64_t)v0;
        p[1] = (int64_t)e0;
        p[2] = (int64_t)s0;
        p[3] = (int64_t)e1;
        if (!(e0_ok && e1_ok)) {  // s0_ok is implied by e1_ok
            if (!e0_ok) p[1] = hpgb_beltidx2nest_fold(rk, flut, ir1, qq[1]);
            if (!s0_ok) p[2] = hpgb_beltidx2nest_fold(rk, flut, ir1 + 1u, qq[2]);
            if (!e1_ok) p[3] = hpgb_beltidx2nest_fold(rk, flut, ir1 + 1u, qq[3]);
        }
    }
#endif
    const double wt = hpgb_div_tame(theta - th[0], th[1] - th[0]);
    w[0] = (1 - w1[0]) * (1 - wt);
    w[1] = w1[0] * (1 - wt);
    w[2] = (1 - w1[1]) * wt;
    w[3] = w1[1] * wt;
}

template <bool NEST>
HPGB_HD void hpgb_interp_cap_lut(const hpgb_geom &g, const hpgb_rk &rk, const double *th_tab, const hpgb_cdiv *cap_tab,
                                 double theta, double phi, uint32_t ir1, int64_t *p, double *w) {
    double th[2], w1[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const uint32_t r = ir1 + (uint32_t)k;
        const bool south = r > rk.ns2;
        const uint32_t nring = south ? rk.ns4 - r : r;
        const double t0 = th_tab[nring];
        th[k] = south ? HPGB_PI - t0 : t0;
        const hpgb_cdiv dphi = cap_tab[nring];
        const uint32_t nr = 4u * nring;
        const int32_t i1 = hpgb_interp_floor32(hpgb_div_const(phi, dphi) - 0.5);
        w1[k] = hpgb_div_const(phi - ((double)i1 + 0.5) * dphi.d, dphi);
        const uint32_t q1 = i1 < 0 ? (uint32_t)i1 + nr : (uint32_t)i1;
        const uint32_t q2 = (uint32_t)(i1 + 1) >= nr ? (uint32_t)(i1 + 1) - nr : (uint32_t)(i1 + 1);
        if (NEST) {
            p[2 * k] = hpgb_capidx2nest_fold(rk, nring, south, q1);
            p[2 * k + 1] = hpgb_capidx2nest_fold(rk, nring, south, q2);
        } else {
            uint32_t sp = 2u * nring * (nring - 1u);
            if (south) sp = rk.npix_lo - sp - nr;
            p[2 * k] = (int64_t)(sp + q1);
            p[2 * k + 1] = (int64_t)(sp + q2);
        }
    }
    const double wt = hpgb_div_tame(theta - th[0], th[1] - th[0]);
    w[0] = (1 - w1[0]) * (1 - wt);
    w[1] = w1[0] * (1 - wt);
    w[2] = (1 - w1[1]) * wt;
    w[3] = w1[1] * wt;
}

// lonlat_to_thetaphi with numpy's formulas (hpgeom/hpgeom.py:75-88): Python-style modulo,
// deg2rad = x * (pi/180).  Returns false when the latitude is out of range.
template <bool DEGREES>
HPGB_HD bool hpgb_lonlat_to_thetaphi_np(double lon, double lat, double &theta, double &phi) {
    const double m = DEGREES ? 360.0 : (2. * HPGB_PI);
    double r = fmod(lon, m);  // numpy's float remainder: result takes the divisor's sign
    if (r != 0.0) {
        if (r < 0.0) r += m;
    } else {
        r = copysign(0.0, m);
    }
    double lon_ = r, lat_ = lat;
    if (DEGREES) {
        lon_ = r * (HPGB_PI / 180.0);
        lat_ = lat * (HPGB_PI / 180.0);
    }
    phi = lon_;
    theta = HPGB_PI / 2. - lat_;
    return !((lat_ < -HPGB_PI / 2.) || (lat_ > HPGB_PI / 2.));
}

#endif  // HPGB_FP_H
