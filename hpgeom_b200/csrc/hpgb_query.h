// hpgb_query.h -- query_circle in the RING scheme, reorganised for the GPU.
//
// The reference (query_disc, hpgeom/healpix_geom.c:632-747) walks the iso-latitude rings the
// disc touches one after the other and appends, per ring, the pixel interval(s) whose centres
// (inclusive: whose area) lie inside; the wrapper then expands the intervals into a pixel list
// (hpgeom/hpgeom.c:678-720).  Every ring is independent, so here
//   * the O(1) per-disc scalars (cosines of the radius, ring limits, the two polar ranges) are
//     prepared once on the host -- hpgb_disc, filled by hpgb_query_circle_ring_plan();
//   * hpgb_qc_ring() below is the reference's loop body for ONE ring, operation for operation
//     (k_query.cu runs it with one thread per ring and writes two rows of an (M,2) range table);
//   * the table is expanded by the pixel_ranges_to_pixels kernels (k_ranges.cu).
// The trigonometry that decides pixel membership (atan2, and cos for the inclusive edge test) is
// the correctly-rounded double-double version of hpgb_math.h, i.e. the value glibc returns
// wherever glibc is itself correctly rounded.
#ifndef HPGB_QUERY_H
#define HPGB_QUERY_H

#include <math.h>

#include "../../include/hpgb.h"
#include "hpgb_core.h"
#include "hpgb_fp.h"

// ring -> z of its pixel centres: hpgeom/healpix_geom.c:456-465
HPGB_HD double hpgb_ring2z(const hpgb_geom &g, int64_t ring) {
    if (ring < g.nside) return 1 - (double)(ring * ring) * g.fact2;
    if (ring <= 3 * g.nside) return (double)(2 * g.nside - ring) * g.fact1;
    ring = 4 * g.nside - ring;
    return (double)(ring * ring) * g.fact2 - 1;
}

// hpgeom/healpix_geom.c:512-528
HPGB_HD void hpgb_ring_info_small(const hpgb_geom &g, int64_t ring, int64_t &startpix, int64_t &ringpix, bool &shifted) {
    if (ring < g.nside) {
        shifted = true;
        ringpix = 4 * ring;
        startpix = 2 * ring * (ring - 1);
    } else if (ring < 3 * g.nside) {
        shifted = ((ring - g.nside) & 1) == 0;
        ringpix = 4 * g.nside;
        startpix = g.ncap + (ring - g.nside) * ringpix;
    } else {
        shifted = true;
        const int64_t nr = 4 * g.nside - ring;
        ringpix = 4 * nr;
        startpix = g.npix - 2 * nr * (nr + 1);
    }
}

// hpgeom/healpix_geom.c:530-532
HPGB_HD double hpgb_cosdist_zphi(const double *T, double z1, double phi1, double z2, double phi2) {
    return z1 * z2 + hpgb_cos_cr(T, phi1 - phi2) * sqrt((1. - z1 * z1) * (1. - z2 * z2));
}

// true = the pixel does NOT overlap the disc (inclusive mode: its edge, sampled at the resolution
// fct * nside, stays outside the enlarged radius): hpgeom/healpix_geom.c:603-630
HPGB_HD bool hpgb_check_pixel_ring(const hpgb_geom &g1, const hpgb_geom &g2, const double *T, int64_t pix, int64_t nr,
                                   int64_t ipix1, int64_t fct, double cz, double cphi, double cosrp2, int64_t cpix) {
    if (pix >= nr) pix -= nr;
    if (pix < 0) pix += nr;
    pix += ipix1;
    if (pix == cpix) return false;  // disc centre in the pixel
    int32_t px, py;
    uint32_t pf;
    hpgb_ring2xyf<false>(g1, pix, px, py, pf);
    const int32_t f = (int32_t)fct, ox = f * px, oy = f * py;
    for (int32_t i = 0; i < f - 1; i++) {  // along the four edges
        double pz, pphi, sth;
        bool hs;
        hpgb_pix2loc<false, false>(g2, hpgb_xyf2ring(g2, ox + i, oy, pf), pz, pphi, sth, hs);
        if (hpgb_cosdist_zphi(T, pz, pphi, cz, cphi) > cosrp2) return false;
        hpgb_pix2loc<false, false>(g2, hpgb_xyf2ring(g2, ox + f - 1, oy + i, pf), pz, pphi, sth, hs);
        if (hpgb_cosdist_zphi(T, pz, pphi, cz, cphi) > cosrp2) return false;
        hpgb_pix2loc<false, false>(g2, hpgb_xyf2ring(g2, ox + f - 1 - i, oy + f - 1, pf), pz, pphi, sth, hs);
        if (hpgb_cosdist_zphi(T, pz, pphi, cz, cphi) > cosrp2) return false;
        hpgb_pix2loc<false, false>(g2, hpgb_xyf2ring(g2, ox, oy + f - 1 - i, pf), pz, pphi, sth, hs);
        if (hpgb_cosdist_zphi(T, pz, pphi, cz, cphi) > cosrp2) return false;
    }
    return true;
}

// One ring of the reference's loop (hpgeom/healpix_geom.c:676-735), in three pieces so that the kernel
// can hand the (rare, long) inclusive edge scans to the whole grid:
//   hpgb_qc_ring_raw  : :677-699, the in-ring index interval [ip_lo, ip_hi] cut out by the cone of
//                       radius rbig (false: the ring contributes nothing);
//   hpgb_qc_ring_trim : :701-710, inclusive mode only: drop pixels from both ends while they do not
//                       overlap the disc -- at most max_steps checks per end; returns a mask of the ends
//                       still unresolved (bit 0: low end, bit 1: high end);
//   hpgb_qc_ring_rows : :712-733, the up-to-two [lo, hi) pixel intervals, ascending; unused row (0, 0).
HPGB_HD bool hpgb_qc_ring_raw(const hpgb_disc_t &d, const hpgb_geom &g, const double *T, int64_t iz, int64_t &ip_lo,
                              int64_t &ip_hi, int64_t &nr, int64_t &ipix1) {
    const double z = hpgb_ring2z(g, iz);
    const double x = (d.cosrbig - z * d.z0) * d.xa;
    const double ysq = 1 - z * z - x * x;
    double dphi;
    if (ysq <= 0)  // no intersection, ring completely inside or outside
        dphi = (d.fct == 1) ? 0 : HPGB_PI - 1e-15;
    else
        dphi = hpgb_atan2_any(T, sqrt(ysq), x);
    if (!(dphi > 0)) return false;
    bool shifted;
    hpgb_ring_info_small(g, iz, ipix1, nr, shifted);
    const double shift = shifted ? 0.5 : 0.;
    const double scale = (double)nr / HPGB_TWOPI;
    ip_lo = hpgb_d2ll(floor(scale * (d.phi - dphi) - shift)) + 1;
    ip_hi = hpgb_d2ll(floor(scale * (d.phi + dphi) - shift));
    return true;
}

HPGB_HD unsigned hpgb_qc_ring_trim(const hpgb_disc_t &d, const hpgb_geom &g, const hpgb_geom &g2, const double *T, int64_t nr,
                                   int64_t ipix1, int64_t &ip_lo, int64_t &ip_hi, int64_t max_steps) {
    int64_t n = 0;
    while (ip_lo <= ip_hi && hpgb_check_pixel_ring(g, g2, T, ip_lo, nr, ipix1, d.fct, d.z0, d.phi, d.cosrsmall, d.cpix)) {
        ++ip_lo;
        if (++n >= max_steps && ip_lo <= ip_hi) return 3u;  // the high end has not been looked at yet
    }
    n = 0;
    while (ip_hi > ip_lo && hpgb_check_pixel_ring(g, g2, T, ip_hi, nr, ipix1, d.fct, d.z0, d.phi, d.cosrsmall, d.cpix)) {
        --ip_hi;
        if (++n >= max_steps && ip_hi > ip_lo) return 2u;
    }
    return 0u;
}

HPGB_HD void hpgb_qc_ring_rows(int64_t ip_lo, int64_t ip_hi, int64_t nr, int64_t ipix1, int64_t row[4]) {
    row[0] = row[1] = row[2] = row[3] = 0;
    if (ip_lo > ip_hi) return;
    if (ip_hi >= nr) {
        ip_lo -= nr;
        ip_hi -= nr;
    }
    if (ip_lo < 0) {
        row[0] = ipix1;
        row[1] = ipix1 + ip_hi + 1;
        row[2] = ipix1 + ip_lo + nr;
        row[3] = ipix1 + nr;  // ipix2 + 1
        // the reference's range set merges an interval that starts inside the previous one
        // (hpgeom/hpgeom_stack.c:239-243): same union, no pixel twice
        if (row[2] < row[1]) row[2] = row[1];
        if (row[3] < row[2]) row[3] = row[2];
    } else {
        row[0] = ipix1 + ip_lo;
        row[1] = ipix1 + ip_hi + 1;
    }
}

// the plain composition (host simulation, tests)
HPGB_HD void hpgb_qc_ring(const hpgb_disc_t &d, const hpgb_geom &g, const hpgb_geom &g2, const double *T, int64_t iz,
                          int64_t row[4]) {
    row[0] = row[1] = row[2] = row[3] = 0;
    int64_t ip_lo, ip_hi, nr, ipix1;
    if (!hpgb_qc_ring_raw(d, g, T, iz, ip_lo, ip_hi, nr, ipix1)) return;
    if (d.fct > 1) hpgb_qc_ring_trim(d, g, g2, T, nr, ipix1, ip_lo, ip_hi, INT64_MAX);
    hpgb_qc_ring_rows(ip_lo, ip_hi, nr, ipix1, row);
}

// ---- host side: the per-disc plan (libm, like the reference) ----
// max_pixrad of the reference (hpgeom/healpix_geom.c:480-502), host libm: a per-disc scalar
static inline double max_pixrad_host(int64_t nside) {
    const double vaz = 2. / 3.;
    const double phi_a = HPGB_PI / (4. * (double)nside);
    double sintheta = sqrt((1.0 - vaz) * (1.0 + vaz));
    const double vax = sintheta * cos(phi_a);
    const double vay = sintheta * sin(phi_a);
    double t1 = 1. - 1. / (double)nside;
    t1 *= t1;
    const double vbz = 1. - t1 / 3.;
    sintheta = sqrt((1.0 - vbz) * (1.0 + vbz));
    const double vbx = sintheta, vby = 0.0;
    const double cx = vay * vbz - vaz * vby;
    const double cy = vaz * vbx - vax * vbz;
    const double cz = vax * vby - vay * vbx;
    const double vdot = vax * vbx + vay * vby + vaz * vbz;
    return atan2(sqrt(cx * cx + cy * cy + cz * cz), vdot);
}

// hpgeom/healpix_geom.c:638-674 and :737-743, the part of query_disc outside the ring loop
static inline int hpgb_qc_plan(int64_t nside, double theta, double phi, double radius, int64_t fact, hpgb_disc_t *p) {
    if (!p) return HPGB_E_BAD_ARGUMENT;
    const int c = hpgb_check_nside(nside, 0);
    if (c) return c;
    if (fact < 0 || (fact > 0 && fact * nside > ((int64_t)1 << 29))) return HPGB_E_BAD_ARGUMENT;
    const hpgb_geom g = hpgb_make_geom(nside, 0);
    const bool inclusive = fact != 0;
    const int64_t fct = inclusive ? fact : 1;
    *p = hpgb_disc_t();
    p->nside = nside;
    p->fct = fct;
    p->phi = phi;
    p->south_lo = -1;
    p->irmin = 1;
    p->irmax = 0;
    double rsmall, rbig;
    if (fct > 1) {
        rsmall = radius + max_pixrad_host(fct * nside);
        rbig = radius + max_pixrad_host(nside);
    } else {
        rsmall = rbig = inclusive ? radius + max_pixrad_host(nside) : radius;
    }
    if (rsmall >= HPGB_PI) {
        p->whole_sphere = 1;
        p->m = 1;
        return HPGB_OK;
    }
    if (rbig > HPGB_PI) rbig = HPGB_PI;
    p->cosrsmall = cos(rsmall);
    p->cosrbig = cos(rbig);
    const double z0 = cos(theta);
    p->z0 = z0;
    p->xa = 1. / sqrt((1 - z0) * (1 + z0));
    p->cpix = hpgb_loc2pix_exact<false>(g, z0, phi, 0., false);
    const double rlat1 = theta - rsmall;
    const double zmax = cos(rlat1);
    int64_t irmin = hpgb_ring_above(g, zmax) + 1;
    if ((rlat1 <= 0) && (irmin > 1)) {  // north pole in the disc
        int64_t sp, rp;
        bool dummy;
        hpgb_ring_info_small(g, irmin - 1, sp, rp, dummy);
        p->north_hi = sp + rp;
    }
    if ((fct > 1) && (rlat1 > 0)) irmin = irmin - 1 > 1 ? irmin - 1 : 1;
    const double rlat2 = theta + rsmall;
    const double zmin = cos(rlat2);
    int64_t irmax = hpgb_ring_above(g, zmin);
    if ((fct > 1) && (rlat2 < HPGB_PI)) irmax = irmax + 1 < 4 * nside - 1 ? irmax + 1 : 4 * nside - 1;
    if ((rlat2 >= HPGB_PI) && (irmax + 1 < 4 * nside)) {  // south pole in the disc
        int64_t sp, rp;
        bool dummy;
        hpgb_ring_info_small(g, irmax + 1, sp, rp, dummy);
        p->south_lo = sp;
    }
    p->irmin = irmin;
    p->irmax = irmax;
    const int64_t nring = irmax >= irmin ? irmax - irmin + 1 : 0;
    p->m = 2 + 2 * nring;
    return HPGB_OK;
}


// ---------------------------------------------------------------------------------------------
// query_circle in the NEST scheme: the reference's hierarchical search (query_disc, NEST branch,
// hpgeom/healpix_geom.c:748-800, and check_pixel_nest, :541-601).  The reference walks the pixel
// tree depth first with an explicit stack; the verdict on a pixel (drop it / output it or its
// order-`order` ancestor / look at its four children) only depends on the pixel itself, so here the
// tree is walked breadth first, one kernel launch per order over the whole frontier (k_query.cu),
// and the emitted [lo, hi) ranges are sorted afterwards.  The depth-first walk abandons the rest of a
// sub-tree once its ancestor has been output ("unwind the stack"); breadth first that shows up as the
// same single-pixel range emitted more than once, removed after the sort.
// ---------------------------------------------------------------------------------------------
#define HPGB_QN_MAX_ORDER 29
struct hpgb_qn_plan {
    int order, omax, inclusive, whole_sphere;
    double ptg_z, ptg_phi, cosrad;
    double crpdr[HPGB_QN_MAX_ORDER + 1], crmdr[HPGB_QN_MAX_ORDER + 1];
};

// hpgeom/healpix_geom.c:748-771 (host, libm); fact = 0: not inclusive
// dr_cache: optional array of HPGB_QN_MAX_ORDER + 1 doubles, all < 0 before the first call -- the safety distance
// max_pixrad(order) does not depend on the disc, so a batch computes it once
static inline int hpgb_qn_make_plan(int64_t nside, double theta, double phi, double radius, int64_t fact, hpgb_qn_plan *p,
                                    double *dr_cache = nullptr) {
    const int c = hpgb_check_nside(nside, 1);
    if (c) return c;
    if (fact < 0 || (fact & (fact - 1)) || (fact > 0 && fact * nside > ((int64_t)1 << 29))) return HPGB_E_BAD_FACT;
    *p = hpgb_qn_plan();
    p->order = hpgb_ilog2_u64((uint64_t)nside);
    p->inclusive = fact != 0;
    p->omax = p->order + (p->inclusive ? hpgb_ilog2_u64((uint64_t)fact) : 0);
    p->ptg_phi = phi;
    if (radius >= HPGB_PI) {  // the disc covers the whole sphere
        p->whole_sphere = 1;
        return HPGB_OK;
    }
    p->ptg_z = cos(theta);
    p->cosrad = cos(radius);
    for (int o = 0; o <= p->omax; o++) {
        double dr;  // safety distance
        if (dr_cache) {
            if (dr_cache[o] < 0) dr_cache[o] = max_pixrad_host((int64_t)1 << o);
            dr = dr_cache[o];
        } else {
            dr = max_pixrad_host((int64_t)1 << o);
        }
        p->crpdr[o] = ((radius + dr) > HPGB_PI) ? -1. : cos(radius + dr);
        p->crmdr[o] = ((radius - dr) < 0.) ? 1. : cos(radius - dr);
    }
    return HPGB_OK;
}

// Verdict on pixel `pix` of order o (g = geometry of nside 2^o, NEST): 0 = nothing, 1 = output the range
// [lo, hi) at order `order`, 2 = visit the four children 4 pix .. 4 pix + 3 at order o + 1.
// hpgeom/healpix_geom.c:784-796 + check_pixel_nest (:541-601).
HPGB_HD int hpgb_qn_visit(const hpgb_qn_plan &P, const hpgb_geom &g, const double *T, int o, int64_t pix, int64_t &lo, int64_t &hi) {
    double z, phi, sth;
    bool hs;
    hpgb_pix2loc<true, true>(g, pix, z, phi, sth, hs);
    const double cang = hpgb_cosdist_zphi(T, P.ptg_z, P.ptg_phi, z, phi);
    if (!(cang > P.crpdr[o])) return 0;
    const int zone = (cang < P.cosrad) ? 1 : ((cang <= P.crmdr[o]) ? 2 : 3);
    if (o < P.order) {
        if (zone >= 3) {
            const int sdist = 2 * (P.order - o);
            lo = pix << sdist;
            hi = (pix + 1) << sdist;
            return 1;
        }
        return 2;
    }
    if (o > P.order) {  // implies inclusive
        if (zone >= 2 || o >= P.omax) {
            lo = pix >> (2 * (o - P.order));
            hi = lo + 1;
            return 1;
        }
        return 2;
    }
    if (zone >= 2) {
        lo = pix;
        hi = pix + 1;
        return 1;
    }
    if (P.inclusive) {
        if (P.order < P.omax) return 2;
        lo = pix;
        hi = pix + 1;
        return 1;
    }
    return 0;
}

#endif  // HPGB_QUERY_H
