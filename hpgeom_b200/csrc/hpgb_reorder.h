// hpgb_reorder.h -- minimum-instruction NEST <-> RING cores for nside = 2^k (the nest_to_ring /
// ring_to_nest kernels are bound by the integer ALU pipe, not by HBM, so every instruction
// counts: see DESIGN.md 4.2).  Bit-identical to nest2ring / ring2nest of the reference
// (hpgeom/healpix_geom.c:217-227 = nest2xyf/xyf2ring/ring2xyf/xyf2nest, :382-454, :512-528).
//
// In-face coordinates (ix, iy), ring numbers and in-ring offsets are 32-bit; a pixel number is
// produced by ONE unsigned 32x32+64 multiply-add (IMAD.WIDE.U32) whose three operands are
// selected per zone.  With t = jr - nside (jr = ring, 1-based from the north pole), fc = face & 3:
//
//   NEST -> RING, pix_ring = A * B + C
//     north cap (t < 0)       nr = t + nside         A = nr  B = 2nr - 2 + fc     C = nside-1-iy
//     south cap (t >= 2nside) nr = 3nside - t        A = nr  B = -(2nr + 2 - fc)  C = npix + ix - (nr << 32)
//     belt                    A = t + nside/2 - 1    B = 4nside                   C = 2nside + ro
//                             ro = ((jpll*nside + ix - iy + (t&1) - 1) >> 1) mod 4nside
//   derivation: xyf2ring's jp = (jpll*nr + ix - iy + 1 + kshift)/2 reduces to fc*nr + (nside-1-iy) + 1
//   in the north cap and fc*nr + ix + 1 in the south cap (nr = (nside-1-ix)+(nside-1-iy)+1 resp.
//   ix+iy+1); ncap + t*4nside = 4nside*(t + nside/2 - 1) + 2nside; the south product is taken
//   modulo 2^64 (B = 2^32 - (2nr+2-fc), compensated by the -nr in the high word of C).
//
//   RING -> NEST: belt pixels use the edge-line coordinates of the reference's loc2pix
//     jp = ip + (tmp >> 1), jm = jp + nside - tmp   (ip = in-ring index, tmp = ring - nside)
//     face from (jp >> k, jm >> k), ix = jm mod nside, iy = nside-1 - (jp mod nside)
//   (identical to ring2xyf's ifp / ifm, hpgeom/healpix_geom.c:427-436, and its irt / ipt tail);
//   cap pixels: ring from the reference's isqrt (same fp64 expression), q = 0-based in-ring index
//   counted eastward, fc = q div nr, rem = q mod nr (special_div);
//     north: ix = nside - nr + rem, iy = nside-1 - rem      south: ix = rem, iy = nr-1 - rem
//   The reference's isqrt has no correction step, so for 2*ip-1 >~ 2^50 it can put the first / last
//   pixels of a huge cap ring into the neighbouring ring; whenever the in-ring index comes out
//   outside [0, 4nr) the element is handed to the general 64-bit path (hpgb_ring2xyf_any), which
//   replays the reference's arithmetic literally.
#ifndef HPGB_REORDER_H
#define HPGB_REORDER_H

#include "hpgb_core.h"

struct hpgb_rk {
    uint32_t ns, nsm1, ns2, ns3, ns4, ns4m1;
    uint32_t belt_a0, belt_c0;          // nside/2 - 1, 2 nside  (0, 0 for nside = 1)
    uint32_t npix_lo, npix_hi;          // 12 nside^2
    uint32_t npm1_lo, npm1_hi;          // npix - 1
    uint32_t ncap_lo, ncap_hi;          // 2 nside (nside - 1)
    uint32_t lmask, hmask;              // nside^2 - 1
    uint32_t k, k2;                     // order, 2 * order
    uint32_t fsh_hi, fmul_hi;           // order >= 16: 2k - 32 and 1 << (2k - 32)
    uint32_t ns3p1, dnr;                // 3 nside + 1, nside - (3 nside + 1)
    uint32_t nr_polar;                  // cap rings with nr <= nr_polar have |z| > 0.99 (hpgeom/healpix_geom.c:317,336)
};

HPGB_HD hpgb_rk hpgb_make_rk(const hpgb_geom &g) {
    hpgb_rk c;
    const uint32_t ns = (uint32_t)g.nside;
    const uint32_t k = (uint32_t)(g.order < 0 ? 0 : g.order);
    c.ns = ns;
    c.nsm1 = ns - 1u;
    c.ns2 = 2u * ns;
    c.ns3 = 3u * ns;
    c.ns4 = 4u * ns;
    c.ns4m1 = 4u * ns - 1u;
    c.belt_a0 = k >= 1 ? ns / 2u - 1u : 0u;
    c.belt_c0 = k >= 1 ? 2u * ns : 0u;
    c.npix_lo = (uint32_t)(uint64_t)g.npix;
    c.npix_hi = (uint32_t)((uint64_t)g.npix >> 32);
    c.npm1_lo = (uint32_t)(uint64_t)(g.npix - 1);
    c.npm1_hi = (uint32_t)((uint64_t)(g.npix - 1) >> 32);
    c.ncap_lo = (uint32_t)(uint64_t)g.ncap;
    c.ncap_hi = (uint32_t)((uint64_t)g.ncap >> 32);
    c.lmask = (uint32_t)(uint64_t)(g.npface - 1);
    c.hmask = (uint32_t)((uint64_t)(g.npface - 1) >> 32);
    c.k = k;
    c.k2 = 2u * k;
    c.fsh_hi = k >= 16 ? 2u * k - 32u : 0u;
    c.fmul_hi = k >= 16 ? (1u << (2u * k - 32u)) : 0u;
    c.ns3p1 = 3u * ns + 1u;
    c.dnr = ns - (3u * ns + 1u);
    // |z| = 1 - nr^2 fact2 decreases with nr: the largest nr for which the reference's test z > 0.99 holds,
    // evaluated with the reference's own expression
    uint32_t lo = 0u, hi = ns;  // P(lo) true (z = 1), hi exclusive
    while (hi - lo > 1u) {
        const uint32_t mid = lo + (hi - lo) / 2u;
        const double t = ((double)mid * (double)mid) * g.fact2;
        if (1.0 - t > 0.99) lo = mid; else hi = mid;
    }
    c.nr_polar = lo;
    return c;
}

// jpll[face] = {1,3,5,7,0,2,4,6,1,3,5,7} as a byte lookup done by one PRMT: selector nibble 0 =
// face & 7 picks the entry, nibbles 1-3 = 4 pick the zero byte.
HPGB_HD uint32_t hpgb_jpll_lut(uint32_t face) { return hpgb_prmt(0x07050301u, 0x06040200u, (face & 7u) | 0x4440u); }

// (face, ix, iy) -> RING pixel (xyf2ring of the reference, hpgeom/healpix_geom.c:394-408, in the
// three-zone multiply-add form described at the top of this file)
HPGB_HD int64_t hpgb_xyf2ring_k(const hpgb_rk &c, uint32_t ix, uint32_t iy, uint32_t face) {
#if defined(__CUDA_ARCH__)
    // Zone selection written as predicated PTX so that it stays ~20 instructions: belt operands by
    // default, overwritten under the north / south predicates (the C++ below is the same arithmetic;
    // left to the compiler it becomes either a divergent branch per element or ~35 selects).
    uint64_t res;
    asm("{\n\t"
        ".reg .pred pn, ps;\n\t"
        ".reg .u32 fc, fr, fm2, frns, t, v, d, w, ks, X, A, B, Cl, Ch, niy;\n\t"
        ".reg .u64 C;\n\t"
        "and.b32 fc, %3, 3;\n\t"
        "shr.u32 fr, %3, 2;\n\t"
        "mul.lo.u32 frns, fr, %4;\n\t"
        "sub.u32 t, %5, %1;\n\t"
        "sub.u32 t, t, %2;\n\t"
        "add.u32 t, t, frns;\n\t"          // t = fr*ns + (ns-1) - ix - iy = ring - nside
        "setp.lt.s32 pn, t, 0;\n\t"        // north cap
        "setp.ge.s32 ps, t, %6;\n\t"       // south cap (t >= 2 nside)
        "and.b32 v, frns, %4;\n\t"         // (fr & 1) * nside
        "sub.u32 d, %1, %2;\n\t"
        "mad.lo.u32 w, fc, %6, d;\n\t"     // fc * 2 nside + ix - iy
        "and.b32 ks, t, 1;\n\t"
        "add.u32 X, w, ks;\n\t"
        "add.u32 X, X, %5;\n\t"
        "sub.u32 X, X, v;\n\t"             // jpll * nside + ix - iy + kshift - 1
        "shr.u32 X, X, 1;\n\t"
        "and.b32 X, X, %9;\n\t"
        "add.u32 Cl, X, %11;\n\t"          // belt: 2 nside + in-ring offset
        "add.u32 A, t, %10;\n\t"           // belt: t + nside/2 - 1
        "mov.u32 B, %8;\n\t"               // belt: 4 nside
        "mov.u32 Ch, 0;\n\t"
        "sub.u32 fm2, fc, 2;\n\t"
        "@pn add.u32 A, t, %4;\n\t"        // north: nr = t + nside
        "@ps sub.u32 A, %7, t;\n\t"        // south: nr = 3 nside - t
        "@pn mad.lo.u32 B, A, 2, fm2;\n\t"
        "@ps mad.lo.u32 B, A, 0xfffffffe, fm2;\n\t"
        "not.b32 niy, %2;\n\t"
        "@pn and.b32 Cl, niy, %5;\n\t"     // north: nside - 1 - iy
        "@ps or.b32 Cl, %12, %1;\n\t"      // south: npix_lo + ix
        "@ps sub.u32 Ch, %13, A;\n\t"      // south: npix_hi - nr (compensates B = 2^32 - ...)
        "mov.b64 C, {Cl, Ch};\n\t"
        "mad.wide.u32 %0, A, B, C;\n\t"
        "}"
        : "=l"(res)
        : "r"(ix), "r"(iy), "r"(face), "r"(c.ns), "r"(c.nsm1), "r"(c.ns2), "r"(c.ns3), "r"(c.ns4), "r"(c.ns4m1),
          "r"(c.belt_a0), "r"(c.belt_c0), "r"(c.npix_lo), "r"(c.npix_hi));
    return (int64_t)res;
#else
    const uint32_t fr = face >> 2, fc = face & 3u;
    const uint32_t t = fr * c.ns + c.nsm1 - ix - iy;  // jr - nside, two's complement
    const bool north = (int32_t)t < 0, south = (int32_t)t >= (int32_t)c.ns2;
    // jpll = 2 fc + 1 - (fr & 1)
    const uint32_t X = fc * c.ns2 + (ix - iy) + (t & 1u) + c.nsm1 - ((fr * c.ns) & c.ns);
    uint32_t A = t + c.belt_a0, B = c.ns4, Cl = ((X >> 1) & c.ns4m1) + c.belt_c0, Ch = 0u;
    const uint32_t fm2 = fc - 2u;
    if (north) {
        A = t + c.ns;
        B = 2u * A + fm2;
        Cl = ~iy & c.nsm1;
    } else if (south) {
        A = c.ns3 - t;
        B = fm2 - 2u * A;
        Cl = c.npix_lo | ix;
        Ch = c.npix_hi - A;
    }
    return (int64_t)((uint64_t)A * (uint64_t)B + (((uint64_t)Ch << 32) | Cl));
#endif
}

// NEST pixel (lo, hi words; must be < npix) -> RING pixel.  HI: order >= 16.
template <bool HI>
HPGB_HD int64_t hpgb_nest2ring_k(const hpgb_rk &c, uint32_t lo, uint32_t hi) {
    const uint32_t face = HI ? (hi >> c.fsh_hi) : (uint32_t)((((uint64_t)hi << 32) | lo) >> c.k2);
    uint32_t wl = HI ? lo : (lo & c.lmask), wh = hi & c.hmask;
    HPGB_DSWAP(wl, 0x22222222u, 1);
    HPGB_DSWAP(wh, 0x22222222u, 1);
    HPGB_DSWAP(wl, 0x0C0C0C0Cu, 2);
    HPGB_DSWAP(wh, 0x0C0C0C0Cu, 2);
    HPGB_DSWAP(wl, 0x00F000F0u, 4);
    HPGB_DSWAP(wh, 0x00F000F0u, 4);
    return hpgb_xyf2ring_k(c, hpgb_prmt(wl, wh, 0x6420), hpgb_prmt(wl, wh, 0x7531), face);
}

// ---- RING -> NEST, in pieces (the streaming kernel runs the belt and the cap part in different
// phases, see OpReorder::tile in k_reorder.cu; hpgb_ring2nest_k below is the plain composition) ----

// ring - nside of a valid RING pixel (two's complement: negative in the north cap, > 2 nside in the
// south cap) and the low word of its index inside the belt
HPGB_HD uint32_t hpgb_r2n_zone(const hpgb_rk &c, uint64_t up, uint32_t &ipb_lo) {
    const uint64_t ipb = up - (((uint64_t)c.ncap_hi << 32) | c.ncap_lo);
    ipb_lo = (uint32_t)ipb;
    return (uint32_t)((int64_t)ipb >> (c.k + 2u));  // fits int32
}
HPGB_HD bool hpgb_r2n_is_cap(const hpgb_rk &c, uint32_t tmp) { return tmp > c.ns2; }

HPGB_HD void hpgb_r2n_belt(const hpgb_rk &c, uint32_t tmp, uint32_t ipb_lo, uint32_t &ix, uint32_t &iy, uint32_t &face) {
    const uint32_t ip = ipb_lo & c.ns4m1;
    const uint32_t jp = ip + (tmp >> 1);
    const uint32_t jm = jp + c.ns - tmp;
    const uint32_t ifp = jp >> c.k, ifm = jm >> c.k;
    face = (ifp == ifm) ? (ifp | 4u) : ((ifp < ifm) ? ifp : (ifm + 8u));
    ix = jm & c.nsm1;
    iy = ~jp & c.nsm1;
}

// polar-cap pixel.  Everything after the reference's isqrt is 32-bit: the ring number is < 2^30
// and the in-ring index is compared modulo 2^32, which still detects a ring that is off by one
// (|error| of the uncorrected isqrt <= 1, and 4 * ring <= 2^31).
HPGB_HD void hpgb_r2n_cap(const hpgb_rk &c, const hpgb_geom &g, uint64_t up, bool north, uint32_t &ix, uint32_t &iy,
                          uint32_t &face) {
    const uint64_t ipm1 = north ? up : ((((uint64_t)c.npm1_hi << 32) | c.npm1_lo) - up);  // 0-based from the pole
    // isqrt of the reference: (int64)sqrt((double)(2 ip - 1) + 0.5), hpgeom/healpix_geom.c:60; < 2^31
#if defined(__CUDA_ARCH__)
    const uint32_t isq = __double2uint_rz(sqrt((double)(int64_t)(2 * ipm1 + 1) + 0.5));
#else
    const uint32_t isq = (uint32_t)(int64_t)sqrt((double)(int64_t)(2 * ipm1 + 1) + 0.5);
#endif
    const uint32_t ir = (isq + 1u) >> 1;
    const uint32_t q0 = (uint32_t)ipm1 - ir * (2u * ir - 2u);  // in-ring index, eastward, counted from the pole side
    if (q0 >= 4u * ir) {  // the reference's isqrt mis-rounded: replay its 64-bit arithmetic
        int32_t sx, sy;
        hpgb_ring2xyf_any(g, (int64_t)up, sx, sy, face);
        ix = (uint32_t)sx;
        iy = (uint32_t)sy;
        return;
    }
    uint32_t q = north ? q0 : 4u * ir - 1u - q0;
    const uint32_t t1 = (q >= 2u * ir) ? 1u : 0u;
    q -= t1 ? 2u * ir : 0u;
    const uint32_t t2 = (q >= ir) ? 1u : 0u;
    q -= t2 ? ir : 0u;
    face = 2u * t1 + t2 + (north ? 0u : 8u);
    ix = north ? c.ns - ir + q : q;
    iy = north ? c.nsm1 - q : ir - 1u - q;
}

// ---- RING -> NEST through FOLDED coordinates (the streaming kernel, OpReorder::tile in k_reorder.cu) ----
// nest = (face << 2k) + interleave(ix, iy): the four face bits land on bits 2k .. 2k+3 of the pixel number, i.e. on
// bits k, k+1 of the even (ix) and of the odd (iy) bit plane.  So with
//     X = ix + ((f0 | f2 << 1) << k),   Y = iy + ((f1 | f3 << 1) << k)        (X, Y < 2^(k+2) <= 2^31)
// the pixel number is interleave(X, Y): no face arithmetic after the Morton step, the pair (X, Y) is the 8-byte
// intermediate a shared-memory slot holds between the passes, and ANY 64-bit result v (-1 for a refused pixel, the
// reference's garbage for its isqrt-quirk pixels) can be stored as deinterleave(v).
// Belt pixels: the face of the reference's ring2xyf depends only on (ifp, ifm) = (jp >> k, jm >> k), ifm - ifp in
// {-1, 0, 1}, ifp in 0..4 -> a 16-entry table indexed by 2 ifp + ifm + 1 (128 bytes of shared memory, one entry per
// pair of banks: conflict-free LDS.64) that holds the two constants of
//     X = jm + lx,  lx = (fx - ifm) << k          Y = ly - jp,  ly = ((fy + ifp + 1) << k) - 1
// (jm = ifm nside + ix, jp = ifp nside + (nside-1-iy)): one load and two adds replace the face selection, the two
// masks and the face multiply-add.
struct alignas(8) hpgb_u32x2 {
    uint32_t x, y;
};
HPGB_HD void hpgb_r2n_lut_entry(const hpgb_rk &c, uint32_t idx, uint32_t &lx, uint32_t &ly) {
    // idx = 2 ifp + ifm + 1 with ifm = ifp + d - 1, d in 0..2  ->  ifp = idx / 3, d = idx % 3
    const uint32_t ifp = idx / 3u, ifm = ifp + idx % 3u - 1u;
    const uint32_t face = (ifp == ifm) ? (ifp | 4u) : ((ifp < ifm) ? ifp : (ifm + 8u));
    const uint32_t fx = (face & 1u) | ((face >> 1) & 2u), fy = ((face >> 1) & 1u) | ((face >> 2) & 2u);
    lx = (fx - ifm) << c.k;
    ly = ((fy + ifp + 1u) << c.k) - 1u;
}
// belt pixel -> (X, Y); lut = the 16 (lx, ly) pairs
HPGB_HD void hpgb_r2n_belt_fold(const hpgb_rk &c, const hpgb_u32x2 *lut, uint32_t tmp, uint32_t ipb_lo, uint32_t &X, uint32_t &Y) {
    const uint32_t ip = ipb_lo & c.ns4m1;
    const uint32_t jp = ip + (tmp >> 1);
    const uint32_t jm = jp + c.ns - tmp;
    const hpgb_u32x2 f = lut[2u * (jp >> c.k) + (jm >> c.k) + 1u];
    X = jm + f.x;
    Y = f.y - jp;
}
// cap pixel -> (X, Y) (hpgb_r2n_cap with the face folded in; the quirk pixels keep the literal 64-bit replay)
template <bool HI>
HPGB_HD void hpgb_r2n_cap_fold(const hpgb_rk &c, const hpgb_geom &g, uint64_t up, uint32_t &X, uint32_t &Y);

// RING pixel (must be < npix) -> (ix, iy, face): ring2xyf of the reference (hpgeom/healpix_geom.c:410-454)
// from the pieces above
HPGB_HD void hpgb_ring2xyf_k(const hpgb_rk &c, const hpgb_geom &g, int64_t pix, uint32_t &ix, uint32_t &iy, uint32_t &face) {
    uint32_t ipb_lo;
    const uint32_t tmp = hpgb_r2n_zone(c, (uint64_t)pix, ipb_lo);
    if (hpgb_r2n_is_cap(c, tmp))
        hpgb_r2n_cap(c, g, (uint64_t)pix, (int32_t)tmp < 0, ix, iy, face);
    else
        hpgb_r2n_belt(c, tmp, ipb_lo, ix, iy, face);
}

// upgrade_pixels in the RING scheme without any Morton code (reference chain: ring_to_nest -> children
// p << 2k .. -> nest_to_ring, hpgeom/hpgeom.py:515-556): parent -> (ix, iy, face); its children are the
// 2^k x 2^k block (ix << k | cx, iy << k | cy) of the same face, in nest order (child number =
// interleave(cx, cy)); this converts the C consecutive children starting at child number `first`
// (a multiple of C; C = 1 or 4).  `last` = the parent is nest pixel npix-1 (the reference refuses it).
template <int C>
HPGB_HD void hpgb_upgrade_ring_group(const hpgb_rk &c, const hpgb_rk &cu, const hpgb_geom &g, int64_t p, uint64_t first,
                                     int64_t *r, bool &last) {
    uint32_t ix, iy, face, gx, gy;
    hpgb_ring2xyf_k(c, g, p, ix, iy, face);
    last = face == 11u && ix == c.nsm1 && iy == c.nsm1;
    const uint32_t kk = cu.k - c.k;
    gx = gy = 0u;
    if (first != 0) hpgb_deinterleave(first, gx, gy);  // nside_upgrade = 2 nside: always the first (only) group
    const uint32_t bx = (ix << kk) | gx, by = (iy << kk) | gy;
#pragma unroll
    for (int j = 0; j < C; j++) r[j] = hpgb_xyf2ring_k(cu, bx + (uint32_t)(j & 1), by + (uint32_t)(j >> 1), face);
}

template <bool HI>
HPGB_HD int64_t hpgb_xyf2nest_k(const hpgb_rk &c, uint32_t ix, uint32_t iy, uint32_t face) {
    uint32_t lo = hpgb_prmt(ix, iy, 0x5140), hi = hpgb_prmt(ix, iy, 0x7362);
    HPGB_DSWAP(lo, 0x00F000F0u, 4);
    HPGB_DSWAP(hi, 0x00F000F0u, 4);
    HPGB_DSWAP(lo, 0x0C0C0C0Cu, 2);
    HPGB_DSWAP(hi, 0x0C0C0C0Cu, 2);
    HPGB_DSWAP(lo, 0x22222222u, 1);
    HPGB_DSWAP(hi, 0x22222222u, 1);
    if (HI) return (int64_t)(((uint64_t)(hi + face * c.fmul_hi) << 32) | lo);
    return (int64_t)((((uint64_t)hi << 32) | lo) + ((uint64_t)face << c.k2));
}

template <bool HI>
HPGB_HD void hpgb_r2n_cap_fold(const hpgb_rk &c, const hpgb_geom &g, uint64_t up, uint32_t &X, uint32_t &Y) {
    // a cap pixel is north iff it lies in the first half of the map (HI: npix / 2 is a multiple of 2^32)
    const bool north = HI ? ((uint32_t)(up >> 32) < (c.npix_hi >> 1)) : (up < (((uint64_t)c.ncap_hi << 32) | c.ncap_lo));
    const uint64_t ipm1 = north ? up : ((((uint64_t)c.npm1_hi << 32) | c.npm1_lo) - up);  // 0-based from the pole
#if defined(__CUDA_ARCH__)
    const uint32_t isq = __double2uint_rz(sqrt((double)(int64_t)(2 * ipm1 + 1) + 0.5));  // hpgeom/healpix_geom.c:60
#else
    const uint32_t isq = (uint32_t)(int64_t)sqrt((double)(int64_t)(2 * ipm1 + 1) + 0.5);
#endif
    const uint32_t ir = (isq + 1u) >> 1;
    const uint32_t q0 = (uint32_t)ipm1 - ir * (2u * ir - 2u);
    if (q0 >= 4u * ir) {  // the reference's isqrt mis-rounded: replay its 64-bit arithmetic, store the result de-interleaved
        int32_t sx, sy;
        uint32_t face;
        hpgb_ring2xyf_any(g, (int64_t)up, sx, sy, face);
        hpgb_deinterleave((uint64_t)hpgb_xyf2nest_k<HI>(c, (uint32_t)sx, (uint32_t)sy, face), X, Y);
        return;
    }
    // face = 2 t1 + t2 (+ 8 south): f0 = t2 rides on X, f1 = t1 and f3 = south on Y
    uint32_t q = north ? q0 : 4u * ir - 1u - q0;
    const bool t1 = q >= 2u * ir;
    q -= t1 ? 2u * ir : 0u;
    const bool t2 = q >= ir;
    q -= t2 ? ir : 0u;
    const uint32_t fy = (t1 ? c.ns : 0u) + (north ? 0u : c.ns2);
    X = (north ? c.ns - ir + q : q) + (t2 ? c.ns : 0u);
    Y = (north ? c.nsm1 - q : ir - 1u - q) + fy;
}

// (ring number 1..4 nside-1, 0-based in-ring index counted eastward) -> NEST pixel: ring2xyf without
// the isqrt, for callers that already know the ring (interpolation): caps fc = q div nr (special_div),
// belt the edge-line coordinates.  Branch-free.
HPGB_HD void hpgb_ringidx2xyf_k(const hpgb_rk &c, uint32_t ring, uint32_t q, uint32_t &ix, uint32_t &iy, uint32_t &face) {
    const bool south = ring > c.ns2;
    const uint32_t nring = south ? c.ns4 - ring : ring;
    const bool cap = nring < c.ns;
    // caps
    const uint32_t nr = cap ? nring : c.ns;
    uint32_t r = q;
    const uint32_t t1 = (r >= 2u * nr) ? 1u : 0u;
    r -= t1 ? 2u * nr : 0u;
    const uint32_t t2 = (r >= nr) ? 1u : 0u;
    r -= t2 ? nr : 0u;
    const uint32_t face_c = 2u * t1 + t2 + (south ? 8u : 0u);
    const uint32_t ix_c = south ? r : c.ns - nr + r;
    const uint32_t iy_c = south ? nr - 1u - r : c.nsm1 - r;
    // belt
    const uint32_t tmp = ring - c.ns;
    const uint32_t jp = q + (tmp >> 1);
    const uint32_t jm = jp + c.ns - tmp;
    const uint32_t ifp = jp >> c.k, ifm = jm >> c.k;
    const uint32_t face_b = hpgb_sel(ifp == ifm, ifp | 4u, hpgb_sel(ifp < ifm, ifp, ifm + 8u));
    ix = hpgb_sel(cap, ix_c, jm & c.nsm1);
    iy = hpgb_sel(cap, iy_c, ~jp & c.nsm1);
    face = hpgb_sel(cap, face_c, face_b);
}
HPGB_HD int64_t hpgb_ringidx2nest_k(const hpgb_rk &c, uint32_t ring, uint32_t q) {
    uint32_t ix, iy, face;
    hpgb_ringidx2xyf_k(c, ring, q, ix, iy, face);
    return hpgb_xyf2nest_k<false>(c, ix, iy, face);
}

// Two pixels of one ring (the interpolation pair): q2 is normally q1 + 1, the EAST neighbour, which inside
// a base face is (ix + 1, iy - 1) -- one step on the even and one on the odd bit plane of the first
// pixel's Morton code (dilated-integer arithmetic, a dozen instructions instead of a second conversion).
// Pairs that straddle a face edge or wrap around the ring take the plain conversion (out of line).
HPGB_HD_NOINLINE int64_t hpgb_ringidx2nest_far(const hpgb_rk &c, uint32_t ring, uint32_t q) {
    return hpgb_ringidx2nest_k(c, ring, q);
}
HPGB_HD void hpgb_ringidx2nest_pair_k(const hpgb_rk &c, uint32_t ring, uint32_t q1, uint32_t q2, int64_t &p1, int64_t &p2) {
    uint32_t ix, iy, face;
    hpgb_ringidx2xyf_k(c, ring, q1, ix, iy, face);
    p1 = hpgb_xyf2nest_k<false>(c, ix, iy, face);
    if (q2 == q1 + 1u && ix < c.nsm1 && iy > 0u) {
        const uint64_t m = (((uint64_t)c.hmask << 32) | c.lmask);  // nside^2 - 1
        const uint64_t EV = 0x5555555555555555ull & m, OD = 0xAAAAAAAAAAAAAAAAull & m;
        const uint64_t v = (uint64_t)p1 & m, fb = (uint64_t)p1 - v;
        const uint64_t xp = (((v & EV) | OD) + 1) & EV;  // ix + 1: the carry rides over the odd plane
        const uint64_t ym = ((v & OD) - 2) & OD;         // iy - 1: the borrow rides over the (zero) even plane
        p2 = (int64_t)(fb + (xp | ym));
    } else {
        p2 = hpgb_ringidx2nest_far(c, ring, q2);
    }
}

// RING pixel (must be < npix) -> NEST pixel the way the streaming kernel's tile routine computes it (fold, then interleave)
template <bool HI>
HPGB_HD int64_t hpgb_ring2nest_fold(const hpgb_rk &c, const hpgb_geom &g, const hpgb_u32x2 *lut, int64_t pix) {
    uint32_t ipb_lo, X, Y;
    const uint32_t tmp = hpgb_r2n_zone(c, (uint64_t)pix, ipb_lo);
    if (hpgb_r2n_is_cap(c, tmp))
        hpgb_r2n_cap_fold<HI>(c, g, (uint64_t)pix, X, Y);
    else
        hpgb_r2n_belt_fold(c, lut, tmp, ipb_lo, X, Y);
    return (int64_t)hpgb_interleave(X, Y);
}

// RING pixel (must be < npix) -> NEST pixel.  HI: order >= 16.
template <bool HI>
HPGB_HD int64_t hpgb_ring2nest_k(const hpgb_rk &c, const hpgb_geom &g, int64_t pix) {
    uint32_t ipb_lo, ix, iy, face;
    const uint32_t tmp = hpgb_r2n_zone(c, (uint64_t)pix, ipb_lo);
    if (hpgb_r2n_is_cap(c, tmp))
        hpgb_r2n_cap(c, g, (uint64_t)pix, (int32_t)tmp < 0, ix, iy, face);
    else
        hpgb_r2n_belt(c, tmp, ipb_lo, ix, iy, face);
    return hpgb_xyf2nest_k<HI>(c, ix, iy, face);
}

#endif  // HPGB_REORDER_H
