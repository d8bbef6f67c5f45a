// hpgb_core.h -- per-element integer HEALPix cores for the sm_100a kernels.
//
// Everything here is `__host__ __device__`: the CUDA kernels in hpgb_kernels.cu are the
// product; tests/hostsim compiles the same element functions with g++ so the bit-level logic
// can be checked against the oracle on a machine without a GPU.  No product code path calls
// the host instantiations.
//
// Design notes (not a translation of the reference's C):
//  * coordinates inside a base face (ix, iy < 2^29), ring numbers (< 2^31) and in-ring
//    offsets (<= 2^31) live in 32-bit registers; only pixel numbers and ring start offsets
//    are 64-bit (Blackwell has a 32-bit integer datapath: every int64 op is 2+ instructions);
//  * Morton interleave / de-interleave is a perfect (un)shuffle of two 32-bit words built
//    from one byte permute (PRMT) plus three masked delta-swaps per word -- no lookup table
//    (the reference uses byte tables, hpgeom/healpix_geom.c:37-55,467-478);
//  * the 3-way north-cap / belt / south-cap branch of the ring scheme is evaluated
//    with selects, not divergent branches;
//  * all results are bit-identical to the reference's (hpgeom/healpix_geom.c, cited below).
#ifndef HPGB_CORE_H
#define HPGB_CORE_H

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define HPGB_HD __host__ __device__ __forceinline__
#define HPGB_HDM __host__ __device__ __forceinline__  // member functions
#define HPGB_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define HPGB_HD static inline
#define HPGB_HDM inline
#define HPGB_HD_NOINLINE static
#endif

// Per-nside constants (reference: struct healpix_info, hpgeom/healpix_geom.h:53-62, built at
// hpgeom/healpix_geom.c:134-151).  Computed once on the host for a scalar nside and passed
// by value as a kernel parameter.
struct hpgb_geom {
    int64_t nside, npface, ncap, npix;
    double fact1, fact2;
    int32_t order;  // log2(nside), or -1 when nside is not a power of two (RING only)
    int32_t nest;
};

HPGB_HD int hpgb_ilog2_u64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return 63 - __clzll((long long)v);
#else
    return 63 - __builtin_clzll(v);
#endif
}

// hpgeom/hpgeom_utils.c:29-46; 0 ok, else HPGB_BAD_NSIDE_* (1,2,3)
HPGB_HD int hpgb_check_nside(int64_t nside, int nest) {
    if (nside <= 0) return 1;
    if (nest && (nside & (nside - 1))) return 2;
    if (nside > ((int64_t)1 << 29)) return 3;
    return 0;
}

HPGB_HD hpgb_geom hpgb_make_geom(int64_t nside, int nest) {
    hpgb_geom g;
    g.nside = nside;
    g.nest = nest;
    g.order = (nside & (nside - 1)) ? -1 : hpgb_ilog2_u64((uint64_t)nside);
    g.npface = nside * nside;
    g.ncap = (g.npface - nside) << 1;
    g.npix = 12 * g.npface;
    g.fact2 = 4. / (double)g.npix;
    g.fact1 = (double)(nside << 1) * g.fact2;
    return g;
}

// uint32 / int32 -> double.  With HPGB_MAGIC_CVT the device version splices the integer into the mantissa of 2^52
// and subtracts 2^52 (exact: one fixed-latency DADD on the fp64 pipe instead of an I2F on the quarter-rate
// conversion unit, whose variable latency also costs a scoreboard wait).  Same double either way.
HPGB_HD double hpgb_u2d(uint32_t v) {
#if defined(__CUDA_ARCH__) && defined(HPGB_MAGIC_CVT)
    return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
#else
    return (double)v;
#endif
}
HPGB_HD double hpgb_i2d(int32_t v) {
#if defined(__CUDA_ARCH__) && defined(HPGB_MAGIC_CVT)
    return __hiloint2double(0x43300000, (int)((uint32_t)v ^ 0x80000000u)) - 4503601774854144.0;  // 2^52 + 2^31
#else
    return (double)v;
#endif
}

// ---------------------------------------------------------------------------------------
// byte permute (PRMT on the device)
// ---------------------------------------------------------------------------------------
HPGB_HD uint32_t hpgb_prmt(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    uint64_t both = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t s = (sel >> (4 * i)) & 7u;
        r |= (uint32_t)((both >> (8 * s)) & 0xFFu) << (8 * i);
    }
    return r;
#endif
}

// masked delta swap: exchanges the bit groups selected by m with the groups s bits above
// them.  On the device each swap is exactly SHF, LOP3, SHF, LOP3 (inline lop3 so that the
// compiler's known-bits rewriting cannot inflate it).
HPGB_HD uint32_t hpgb_xor_and(uint32_t a, uint32_t b, uint32_t c) {  // (a ^ b) & c
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    return (a ^ b) & c;
#endif
}
HPGB_HD uint32_t hpgb_xor3(uint32_t a, uint32_t b, uint32_t c) {  // a ^ b ^ c
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    return a ^ b ^ c;
#endif
}
// w >> s.  With HPGB_SHR_MULHI the device version is a multiply-high by 2^(32-s) whose multiplier
// is laundered through an empty asm so ptxas keeps it an IMAD.HI (FMA pipe) instead of strength-
// reducing it back to SHF (ALU pipe): the integer kernels are ALU-pipe bound.
HPGB_HD uint32_t hpgb_shr(uint32_t w, int s) {
#if defined(__CUDA_ARCH__) && defined(HPGB_SHR_MULHI)
    uint32_t mul = 1u << (32 - s);
    asm("" : "+r"(mul));
    return __umulhi(w, mul);
#else
    return w >> s;
#endif
}
#define HPGB_DSWAP(w, m, s)                                       \
    do {                                                          \
        uint32_t t_ = hpgb_xor_and((w), hpgb_shr((w), (s)), (m)); \
        (w) = hpgb_xor3((w), t_, t_ << (s));                      \
    } while (0)

// Morton-interleave two 29-bit coordinates: bit i of ix -> bit 2i, bit i of iy -> bit 2i+1.
// Same function as spread_bits64(ix) + (spread_bits64(iy) << 1), hpgeom/healpix_geom.c:382-385.
HPGB_HD uint64_t hpgb_interleave(uint32_t ix, uint32_t iy) {
    uint32_t lo = hpgb_prmt(ix, iy, 0x5140);  // bytes [y1 x1 y0 x0]
    uint32_t hi = hpgb_prmt(ix, iy, 0x7362);  // bytes [y3 x3 y2 x2]
    HPGB_DSWAP(lo, 0x00F000F0u, 4);
    HPGB_DSWAP(hi, 0x00F000F0u, 4);
    HPGB_DSWAP(lo, 0x0C0C0C0Cu, 2);
    HPGB_DSWAP(hi, 0x0C0C0C0Cu, 2);
    HPGB_DSWAP(lo, 0x22222222u, 1);
    HPGB_DSWAP(hi, 0x22222222u, 1);
    return ((uint64_t)hi << 32) | lo;
}

// inverse of hpgb_interleave (compress_bits64 of v and of v>>1, hpgeom/healpix_geom.c:387-392)
HPGB_HD void hpgb_deinterleave(uint64_t v, uint32_t &ix, uint32_t &iy) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    HPGB_DSWAP(lo, 0x22222222u, 1);
    HPGB_DSWAP(hi, 0x22222222u, 1);
    HPGB_DSWAP(lo, 0x0C0C0C0Cu, 2);
    HPGB_DSWAP(hi, 0x0C0C0C0Cu, 2);
    HPGB_DSWAP(lo, 0x00F000F0u, 4);
    HPGB_DSWAP(hi, 0x00F000F0u, 4);
    ix = hpgb_prmt(lo, hi, 0x6420);
    iy = hpgb_prmt(lo, hi, 0x7531);
}

// select that stays a select (inline selp: the compiler cannot turn it back into a branch)
HPGB_HD uint32_t hpgb_sel(bool p, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\tselp.u32 %0, %1, %2, q;\n\t}" : "=r"(r) : "r"(a), "r"(b), "r"((uint32_t)p));
    return r;
#else
    return p ? a : b;
#endif
}

// jrll[f] = 2 + (f >> 2); jpll[f] = {1,3,5,7,0,2,4,6,1,3,5,7}  (hpgeom/healpix_geom.c:57-58)
HPGB_HD uint32_t hpgb_jrll(uint32_t f) { return 2u + (f >> 2); }
HPGB_HD uint32_t hpgb_jpll(uint32_t f) { return 2u * (f & 3u) + 1u - (((f >> 2) == 1u) ? 1u : 0u); }

// (face, ix, iy) -> NEST pixel, hpgeom/healpix_geom.c:382-385
HPGB_HD int64_t hpgb_xyf2nest(int order, uint32_t ix, uint32_t iy, uint32_t face) {
    return (int64_t)(((uint64_t)face << (2 * order)) + hpgb_interleave(ix, iy));
}

// NEST pixel -> (face, ix, iy), hpgeom/healpix_geom.c:387-392
HPGB_HD void hpgb_nest2xyf(const hpgb_geom &g, int64_t pix, uint32_t &ix, uint32_t &iy,
                           uint32_t &face) {
    face = (uint32_t)((uint64_t)pix >> (2 * g.order));
    hpgb_deinterleave((uint64_t)pix & (uint64_t)(g.npface - 1), ix, iy);
}

// (face, ix, iy) -> RING pixel.  hpgeom/healpix_geom.c:394-408 with get_ring_info_small
// (:512-528) folded in.  ix, iy: in-face coordinates (< nside).  32-bit arithmetic throughout
// except the ring start offset:  pixels before ring jr = c + u*v  (one signed 32x32->64 IMAD)
//   north cap: 0    + (2 jr)      * (jr - 1)
//   belt     : ncap + (2 (jr-ns)) * (2 ns)
//   south cap: npix - (2 nrs)     * (nrs + 1),   nrs = 4 ns - jr
HPGB_HD int64_t hpgb_xyf2ring(const hpgb_geom &g, int32_t ix, int32_t iy, uint32_t face) {
    const uint32_t ns = (uint32_t)g.nside;
    const uint32_t fr = face >> 2, fc = face & 3u;
    const uint32_t jr = (fr + 2u) * ns - (uint32_t)(ix + iy) - 1u;  // 1 .. 4ns-1
    const bool north = jr < ns;
    const bool south = jr >= 3u * ns;
    const uint32_t nrs = 4u * ns - jr;
    const int32_t u = north ? (int32_t)(2u * jr) : (south ? -(int32_t)(2u * nrs) : (int32_t)(2u * (jr - ns)));
    const int32_t v = north ? (int32_t)(jr - 1u) : (south ? (int32_t)(nrs + 1u) : (int32_t)(2u * ns));
    const int64_t c = north ? (int64_t)0 : (south ? g.npix : g.ncap);
    const int64_t n_before = c + (int64_t)u * (int64_t)v;
    const uint32_t nr = north ? jr : (south ? nrs : ns);
    const uint32_t kshift = (north || south) ? 0u : ((jr - ns) & 1u);
    // jp = (jpll*nr + ix - iy + 1 + kshift) / 2: the numerator is even but needs
    // 33 bits, so halve the two parts separately (A = jpll*nr < 2^32, e fits int32).
    const uint32_t A = (2u * fc + 1u - ((fr == 1u) ? 1u : 0u)) * nr;
    const int32_t e = (ix - iy) + 1 + (int32_t)kshift + (int32_t)(A & 1u);
    uint32_t jp = (A >> 1) + (uint32_t)(e >> 1);
    // "if (jp < 1) jp += 4 nside" can only trigger when jpll == 0, i.e. face 4
    if (face == 4u && (int32_t)jp < 1) jp += 4u * ns;
    return n_before + (int64_t)(uint64_t)(jp - 1u);
}

// isqrt of the reference: (int64)sqrt((double)v + 0.5), no correction step
// (hpgeom/healpix_geom.c:60).  IEEE sqrt / conversions are exact on both sides.
HPGB_HD int64_t hpgb_isqrt(int64_t v) { return (int64_t)sqrt((double)v + 0.5); }

// floor(a/b) for 0 <= a < 4b without a division (hpgeom/healpix_geom.c:102-106)
HPGB_HD uint32_t hpgb_special_div(int64_t a, int64_t b) {
    uint32_t t = (a >= (b << 1)) ? 1u : 0u;
    a -= t ? (b << 1) : 0;
    return (t << 1) + ((a >= b) ? 1u : 0u);
}

// RING pixel -> (face, ix, iy), hpgeom/healpix_geom.c:410-454, general nside (RING allows
// non-powers of two): 64-bit arithmetic and true divisions, as the reference.
HPGB_HD_NOINLINE void hpgb_ring2xyf_any(const hpgb_geom &g, int64_t pix, int32_t &ix, int32_t &iy, uint32_t &face) {
    const int64_t ns = g.nside;
    int64_t iring, iphi, nr;
    uint32_t kshift, f;
    if (pix < g.ncap || pix >= g.npix - g.ncap) {  // polar caps (mirror images)
        const bool north = pix < g.ncap;
        const int64_t ip = north ? pix + 1 : g.npix - pix;       // 1-based index from the pole
        const int64_t ir = (1 + hpgb_isqrt(2 * ip - 1)) >> 1;    // ring counted from the pole
        const int64_t off = ip - 2 * ir * (ir - 1);              // 1..4ir from the ring start
        iphi = north ? off : 4 * ir + 1 - off;
        kshift = 0;
        nr = ir;
        f = hpgb_special_div(iphi - 1, nr) + (north ? 0u : 8u);
        iring = north ? ir : 4 * ns - ir;
    } else {  // equatorial belt
        const int64_t ip = pix - g.ncap;
        const int64_t t = ip / (4 * ns);
        iring = t + ns;
        iphi = ip - t * 4 * ns + 1;
        kshift = (uint32_t)((iring + ns) & 1);
        nr = ns;
        const int64_t ire = t + 1, irm = 2 * ns + 1 - t;
        const int64_t ifm = (iphi - (ire >> 1) + ns - 1) / ns;
        const int64_t ifp = (iphi - (irm >> 1) + ns - 1) / ns;
        f = (uint32_t)((ifp == ifm) ? (ifp | 4) : ((ifp < ifm) ? ifp : (ifm + 8)));
    }
    const int64_t irt = iring - (int64_t)(2u + (f >> 2)) * ns + 1;
    int64_t ipt = 2 * iphi - (int64_t)hpgb_jpll(f) * nr - (int64_t)kshift - 1;
    if (ipt >= 2 * ns) ipt -= 8 * ns;
    ix = (int32_t)((ipt - irt) >> 1);
    iy = (int32_t)((-ipt - irt) >> 1);
    face = f;
}

// Same for nside = 2^k: after the ring / in-ring split everything fits 32 bits.  ipt = ix - iy
// is only needed modulo 8 nside (the reference's "if (ipt >= nl2) ipt -= 8*nside"), and 8 nside
// divides 2^32, so it is computed mod 2^32 and sign-extended from bit k+2.
//
// The reference's isqrt has no correction step (hpgeom/healpix_geom.c:60): for 2*ip-1 > 2^53
// the int64 -> double rounding can push the last few pixels of a large cap ring into the next
// ring (in-ring offset < 1 or > 4*ring).  Those pixels must reproduce the reference's int64
// arithmetic exactly, so they take the general path (rare: ~1e-8 of cap pixels at nside 2^29).
HPGB_HD void hpgb_ring2xyf_p2(const hpgb_geom &g, int64_t pix, int32_t &ix, int32_t &iy, uint32_t &face) {
    const uint32_t ns = (uint32_t)g.nside;
    const int k = g.order;
    uint32_t iphi, nr, kshift, f, iring;
    if (pix < g.ncap || pix >= g.npix - g.ncap) {
        const bool north = pix < g.ncap;
        const int64_t ip = north ? pix + 1 : g.npix - pix;
        const uint32_t ir = (uint32_t)((1 + hpgb_isqrt(2 * ip - 1)) >> 1);
        const int64_t off64 = ip - 2 * (int64_t)((uint64_t)ir * (uint64_t)(ir - 1u));
        if (off64 < 1 || off64 > 4 * (int64_t)ir) {
            hpgb_ring2xyf_any(g, pix, ix, iy, face);
            return;
        }
        const uint32_t off = (uint32_t)off64;
        iphi = north ? off : 4u * ir + 1u - off;
        kshift = 0u;
        nr = ir;
        uint32_t a = iphi - 1u;  // special_div(iphi - 1, nr), hpgeom/healpix_geom.c:102-106; 0 <= a < 4 nr
        const uint32_t t = (a >= 2u * ir) ? 1u : 0u;
        a -= t ? 2u * ir : 0u;
        f = 2u * t + ((a >= ir) ? 1u : 0u) + (north ? 0u : 8u);
        iring = north ? ir : 4u * ns - ir;
    } else {
        const uint64_t ip = (uint64_t)(pix - g.ncap);
        const uint32_t t = (uint32_t)(ip >> (k + 2));
        iphi = ((uint32_t)ip & (4u * ns - 1u)) + 1u;
        iring = t + ns;
        kshift = t & 1u;  // (iring + nside) & 1
        nr = ns;
        const uint32_t ifm = (iphi - ((t + 1u) >> 1) + ns - 1u) >> k;
        const uint32_t ifp = (iphi - ((2u * ns + 1u - t) >> 1) + ns - 1u) >> k;
        f = (ifp == ifm) ? (ifp | 4u) : ((ifp < ifm) ? ifp : (ifm + 8u));
    }
    const int32_t irt = (int32_t)(iring - (2u + (f >> 2)) * ns + 1u);
    const uint32_t ipt = 2u * iphi - hpgb_jpll(f) * nr - kshift - 1u;
    const int sh = 29 - k;
    const int32_t ipts = ((int32_t)(ipt << sh)) >> sh;
    ix = (ipts - irt) >> 1;
    iy = (-ipts - irt) >> 1;
    face = f;
}

template <bool POW2>
HPGB_HD void hpgb_ring2xyf(const hpgb_geom &g, int64_t pix, int32_t &ix, int32_t &iy, uint32_t &face) {
    if (POW2)
        hpgb_ring2xyf_p2(g, pix, ix, iy, face);
    else
        hpgb_ring2xyf_any(g, pix, ix, iy, face);
}

// hpgeom/healpix_geom.c:217-227
HPGB_HD int64_t hpgb_nest2ring(const hpgb_geom &g, int64_t pix) {
    uint32_t ix, iy, face;
    hpgb_nest2xyf(g, pix, ix, iy, face);
    return hpgb_xyf2ring(g, (int32_t)ix, (int32_t)iy, face);
}

HPGB_HD int64_t hpgb_ring2nest(const hpgb_geom &g, int64_t pix) {
    int32_t ix, iy;
    uint32_t face;
    hpgb_ring2xyf<true>(g, pix, ix, iy, face);
    return hpgb_xyf2nest(g.order, (uint32_t)ix, (uint32_t)iy, face);
}

HPGB_HD bool hpgb_pixel_ok(const hpgb_geom &g, int64_t pix) {
    return (uint64_t)pix < (uint64_t)g.npix;  // hpgeom/hpgeom_utils.c:62-71
}

#endif  // HPGB_CORE_H
