// hpgb_kernels.cu -- sm_100a kernels + C-ABI entry points (include/hpgb.h) for hpgeom's
// elementwise HEALPix conversion path.  One Op struct per API function (see hpgb_launch.cuh
// for the kernel skeleton they plug into); templated on the flags that the reference tests per
// element (nest / lonlat / degrees, hpgeom/hpgeom.c:143-153) so those branches are
// compile-time here.  A second, scalar instantiation (VAR) serves per-element nside arrays.
#include <cuda_runtime.h>
#include <stdint.h>

#include "hpgb_core.h"
#include "hpgb_fp.h"
#include "hpgb_host.h"
#include "hpgb_launch.cuh"

namespace {

// per-element geometry for nside arrays; flags the element and returns false on a bad nside
// (the reference re-checks nside whenever it changes, hpgeom/hpgeom.c:133-140)
__device__ __forceinline__ bool var_geom(const int64_t *nside_v, int64_t i, int nest, hpgb_geom &g) {
    const int64_t ns = hpgb_ld1(nside_v + i);
    if (hpgb_check_nside(ns, nest)) return false;
    g = hpgb_make_geom(ns, nest);
    return true;
}

// -------------------------------------------------------------------------------------------
// nest_to_ring / ring_to_nest   (reference loops: hpgeom/hpgeom.c:1429-1450, :1699-1720)
// 8 B in + 8 B out per pixel, integer only.
// -------------------------------------------------------------------------------------------
template <bool TO_RING, bool VAR>
struct OpReorder {
    const int64_t *in;
    int64_t *out;
    const int64_t *nside_v;
    int64_t base;  // global index of element 0 (chunked host pipeline)
    hpgb_geom g;
    typedef longlong2 In2;
    typedef hpgb_no_ctx Ctx;

    __device__ __forceinline__ Ctx prologue() const { return Ctx(); }
    __device__ __forceinline__ int64_t conv(const hpgb_geom &gg, int64_t p, int64_t i, hpgb_bad_t &bad) const {
        if (!hpgb_pixel_ok(gg, p)) {
            bad.flag(base + i);
            return -1;
        }
        return TO_RING ? hpgb_nest2ring(gg, p) : hpgb_ring2nest(gg, p);
    }
    __device__ __forceinline__ In2 load2(int64_t i) const { return hpgb_ld2(in + i); }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &, hpgb_bad_t &bad) const {
        hpgb_st2(out + i, conv(g, v.x, i, bad), conv(g, v.y, i + 1, bad));
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &, hpgb_bad_t &bad) const {
        if (VAR) {
            hpgb_geom gg;
            if (!var_geom(nside_v, i, 1, gg)) {
                bad.flag(base + i);
                hpgb_st1(out + i, -1);
                return;
            }
            hpgb_st1(out + i, conv(gg, hpgb_ld1(in + i), i, bad));
        } else {
            hpgb_st1(out + i, conv(g, hpgb_ld1(in + i), i, bad));
        }
    }
};

template <bool TO_RING>
int launch_reorder(const int64_t *in, int64_t *out, int64_t n, int64_t nside, const int64_t *nside_v, int64_t base,
                   hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!in || !out || !st))) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;  // the reference checks nside inside the loop: no elements, no error
    if (nside_v) {
        OpReorder<TO_RING, true> op{in, out, nside_v, base, hpgb_geom()};
        return hpgb_launch_map1(op, n, st, s);
    }
    int c = hpgb_check_nside(nside, 1);
    if (c) return c;
    OpReorder<TO_RING, false> op{in, out, nullptr, base, hpgb_make_geom(nside, 1)};
    return hpgb_launch_map(op, n, hpgb_aligned16(in) && hpgb_aligned16(out), st, s);
}

// -------------------------------------------------------------------------------------------
// angle_to_pixel   (reference loop: hpgeom/hpgeom.c:126-157)   16 B in + 8 B out per point.
// NEST: fast path (hpgb_fp.h) with the exact path as a rare, non-inlined fallback.
// RING: exact path (ring fast path: next).
// -------------------------------------------------------------------------------------------
template <bool NEST, bool LONLAT, bool DEGREES, bool VAR>
struct OpAng2Pix {
    const double *a, *b;
    int64_t *out;
    const int64_t *nside_v;
    int64_t base;
    hpgb_geom g;
    hpgb_a2p_fast f;
    struct In2 {
        double2 a, b;
    };
    typedef const double *Ctx;

    __device__ __forceinline__ Ctx prologue() const { return hpgb_stage_table(); }
    __device__ __forceinline__ In2 load2(int64_t i) const { return In2{hpgb_ld2(a + i), hpgb_ld2(b + i)}; }
    __device__ __forceinline__ int64_t conv(const hpgb_geom &gg, const hpgb_a2p_fast &ff, double va, double vb,
                                            int64_t i, Ctx T, hpgb_bad_t &bad) const {
        int64_t pix;
        if (NEST && hpgb_angle_to_pixel_nest_fast<LONLAT, DEGREES>(gg, ff, va, vb, pix)) return pix;
        if (!hpgb_angle_to_pixel_exact<NEST, LONLAT, DEGREES>(gg, T, va, vb, pix)) bad.flag(base + i);
        return pix;
    }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &T, hpgb_bad_t &bad) const {
        hpgb_st2(out + i, conv(g, f, v.a.x, v.b.x, i, T, bad), conv(g, f, v.a.y, v.b.y, i + 1, T, bad));
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &T, hpgb_bad_t &bad) const {
        if (VAR) {
            hpgb_geom gg;
            if (!var_geom(nside_v, i, NEST, gg)) {
                bad.flag(base + i);
                hpgb_st1(out + i, -1);
                return;
            }
            hpgb_st1(out + i, conv(gg, hpgb_make_a2p_fast(gg), hpgb_ld1(a + i), hpgb_ld1(b + i), i, T, bad));
        } else {
            hpgb_st1(out + i, conv(g, f, hpgb_ld1(a + i), hpgb_ld1(b + i), i, T, bad));
        }
    }
};

template <bool NEST, bool LONLAT, bool DEGREES>
int launch_a2p_t(const double *a, const double *b, int64_t *out, int64_t n, int64_t nside, const int64_t *nside_v,
                 int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (nside_v) {
        OpAng2Pix<NEST, LONLAT, DEGREES, true> op{a, b, out, nside_v, base, hpgb_geom(), hpgb_a2p_fast()};
        return hpgb_launch_map1(op, n, st, s);
    }
    const hpgb_geom g = hpgb_make_geom(nside, NEST);
    OpAng2Pix<NEST, LONLAT, DEGREES, false> op{a, b, out, nullptr, base, g, hpgb_make_a2p_fast(g)};
    return hpgb_launch_map(op, n, hpgb_aligned16(a) && hpgb_aligned16(b) && hpgb_aligned16(out), st, s);
}

int launch_a2p(const double *a, const double *b, int64_t *out, int64_t n, int64_t nside, const int64_t *nside_v,
               int nest, int lonlat, int degrees, int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!a || !b || !out || !st))) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;
    if (!nside_v) {
        int c = hpgb_check_nside(nside, nest);
        if (c) return c;
    }
#define GO(N, L, D) return launch_a2p_t<N, L, D>(a, b, out, n, nside, nside_v, base, st, s)
    if (nest) {
        if (lonlat) {
            if (degrees) GO(true, true, true);
            GO(true, true, false);
        }
        GO(true, false, false);
    }
    if (lonlat) {
        if (degrees) GO(false, true, true);
        GO(false, true, false);
    }
    GO(false, false, false);
#undef GO
}

// -------------------------------------------------------------------------------------------
// pixel_to_angle   (reference loop: hpgeom/hpgeom.c:426-457)   8 B in + 16 B out per pixel
// -------------------------------------------------------------------------------------------
template <bool NEST, bool POW2, bool LONLAT, bool DEGREES, bool VAR>
struct OpPix2Ang {
    const int64_t *pix;
    double *a, *b;
    const int64_t *nside_v;
    int64_t base;
    hpgb_geom g;
    typedef longlong2 In2;
    typedef const double *Ctx;

    __device__ __forceinline__ Ctx prologue() const { return hpgb_stage_table(); }
    __device__ __forceinline__ In2 load2(int64_t i) const { return hpgb_ld2(pix + i); }
    __device__ __forceinline__ void conv(const hpgb_geom &gg, int64_t p, int64_t i, Ctx T, hpgb_bad_t &bad, double &va,
                                         double &vb) const {
        if (!hpgb_pixel_ok(gg, p)) {
            bad.flag(base + i);
            va = vb = 0.0;
            return;
        }
        hpgb_pix2ang<NEST, POW2, LONLAT, DEGREES>(gg, T, p, va, vb);
    }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &T, hpgb_bad_t &bad) const {
        double a0, b0, a1, b1;
        conv(g, v.x, i, T, bad, a0, b0);
        conv(g, v.y, i + 1, T, bad, a1, b1);
        hpgb_st2(a + i, a0, a1);
        hpgb_st2(b + i, b0, b1);
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &T, hpgb_bad_t &bad) const {
        double va = 0.0, vb = 0.0;
        if (VAR) {
            hpgb_geom gg;
            if (!var_geom(nside_v, i, NEST, gg))
                bad.flag(base + i);
            else
                conv(gg, hpgb_ld1(pix + i), i, T, bad, va, vb);
        } else {
            conv(g, hpgb_ld1(pix + i), i, T, bad, va, vb);
        }
        hpgb_st1(a + i, va);
        hpgb_st1(b + i, vb);
    }
};

template <bool NEST, bool LONLAT, bool DEGREES>
int launch_p2a_t(const int64_t *pix, double *a, double *b, int64_t n, int64_t nside, const int64_t *nside_v,
                 int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (nside_v) {  // per-element nside: the general-nside code path handles powers of two as well
        OpPix2Ang<NEST, NEST, LONLAT, DEGREES, true> op{pix, a, b, nside_v, base, hpgb_geom()};
        return hpgb_launch_map1(op, n, st, s);
    }
    const hpgb_geom g = hpgb_make_geom(nside, NEST);
    const bool vec = hpgb_aligned16(pix) && hpgb_aligned16(a) && hpgb_aligned16(b);
    if (NEST || g.order >= 0) {
        OpPix2Ang<NEST, true, LONLAT, DEGREES, false> op{pix, a, b, nullptr, base, g};
        return hpgb_launch_map(op, n, vec, st, s);
    }
    OpPix2Ang<false, false, LONLAT, DEGREES, false> op{pix, a, b, nullptr, base, g};
    return hpgb_launch_map(op, n, vec, st, s);
}

int launch_p2a(const int64_t *pix, double *a, double *b, int64_t n, int64_t nside, const int64_t *nside_v, int nest,
               int lonlat, int degrees, int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!pix || !a || !b || !st))) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;
    if (!nside_v) {
        int c = hpgb_check_nside(nside, nest);
        if (c) return c;
    }
#define GO(N, L, D) return launch_p2a_t<N, L, D>(pix, a, b, n, nside, nside_v, base, st, s)
    if (nest) {
        if (lonlat) {
            if (degrees) GO(true, true, true);
            GO(true, true, false);
        }
        GO(true, false, false);
    }
    if (lonlat) {
        if (degrees) GO(false, true, true);
        GO(false, true, false);
    }
    GO(false, false, false);
#undef GO
}

// -------------------------------------------------------------------------------------------
// vector_to_pixel / pixel_to_vector   (hpgeom/hpgeom.c:2323-2345, :2617-2643)   32 B per element
// -------------------------------------------------------------------------------------------
template <bool NEST, bool VAR>
struct OpVec2Pix {
    const double *x, *y, *z;
    int64_t *out;
    const int64_t *nside_v;
    int64_t base;
    hpgb_geom g;
    struct In2 {
        double2 x, y, z;
    };
    typedef const double *Ctx;
    __device__ __forceinline__ Ctx prologue() const { return hpgb_stage_table(); }
    __device__ __forceinline__ In2 load2(int64_t i) const { return In2{hpgb_ld2(x + i), hpgb_ld2(y + i), hpgb_ld2(z + i)}; }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &T, hpgb_bad_t &) const {
        hpgb_st2(out + i, hpgb_vec2pix<NEST>(g, T, v.x.x, v.y.x, v.z.x), hpgb_vec2pix<NEST>(g, T, v.x.y, v.y.y, v.z.y));
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &T, hpgb_bad_t &bad) const {
        hpgb_geom gg = g;
        if (VAR && !var_geom(nside_v, i, NEST, gg)) {
            bad.flag(base + i);
            hpgb_st1(out + i, -1);
            return;
        }
        hpgb_st1(out + i, hpgb_vec2pix<NEST>(gg, T, hpgb_ld1(x + i), hpgb_ld1(y + i), hpgb_ld1(z + i)));
    }
};

template <bool NEST>
int launch_v2p_t(const double *x, const double *y, const double *z, int64_t *out, int64_t n, int64_t nside,
                 const int64_t *nside_v, int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (nside_v) {
        OpVec2Pix<NEST, true> op{x, y, z, out, nside_v, base, hpgb_geom()};
        return hpgb_launch_map1(op, n, st, s);
    }
    OpVec2Pix<NEST, false> op{x, y, z, out, nullptr, base, hpgb_make_geom(nside, NEST)};
    return hpgb_launch_map(op, n, hpgb_aligned16(x) && hpgb_aligned16(y) && hpgb_aligned16(z) && hpgb_aligned16(out), st, s);
}

int launch_v2p(const double *x, const double *y, const double *z, int64_t *out, int64_t n, int64_t nside,
               const int64_t *nside_v, int nest, int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!x || !y || !z || !out || !st))) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;
    if (!nside_v) {
        int c = hpgb_check_nside(nside, nest);
        if (c) return c;
    }
    return nest ? launch_v2p_t<true>(x, y, z, out, n, nside, nside_v, base, st, s)
                : launch_v2p_t<false>(x, y, z, out, n, nside, nside_v, base, st, s);
}

template <bool NEST, bool POW2, bool VAR>
struct OpPix2Vec {
    const int64_t *pix;
    double *x, *y, *z;
    const int64_t *nside_v;
    int64_t base;
    hpgb_geom g;
    typedef longlong2 In2;
    typedef const double *Ctx;
    __device__ __forceinline__ Ctx prologue() const { return hpgb_stage_table(); }
    __device__ __forceinline__ In2 load2(int64_t i) const { return hpgb_ld2(pix + i); }
    __device__ __forceinline__ void conv(const hpgb_geom &gg, int64_t p, int64_t i, Ctx T, hpgb_bad_t &bad, double &vx,
                                         double &vy, double &vz) const {
        if (!hpgb_pixel_ok(gg, p)) {
            bad.flag(base + i);
            vx = vy = vz = 0.0;
            return;
        }
        hpgb_pix2vec<NEST, POW2>(gg, T, p, vx, vy, vz);
    }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &T, hpgb_bad_t &bad) const {
        double x0, y0, z0, x1, y1, z1;
        conv(g, v.x, i, T, bad, x0, y0, z0);
        conv(g, v.y, i + 1, T, bad, x1, y1, z1);
        hpgb_st2(x + i, x0, x1);
        hpgb_st2(y + i, y0, y1);
        hpgb_st2(z + i, z0, z1);
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &T, hpgb_bad_t &bad) const {
        double vx = 0.0, vy = 0.0, vz = 0.0;
        hpgb_geom gg = g;
        if (VAR && !var_geom(nside_v, i, NEST, gg))
            bad.flag(base + i);
        else
            conv(gg, hpgb_ld1(pix + i), i, T, bad, vx, vy, vz);
        hpgb_st1(x + i, vx);
        hpgb_st1(y + i, vy);
        hpgb_st1(z + i, vz);
    }
};

int launch_p2v(const int64_t *pix, double *x, double *y, double *z, int64_t n, int64_t nside, const int64_t *nside_v,
               int nest, int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!pix || !x || !y || !z || !st))) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;
    if (nside_v) {
        if (nest) {
            OpPix2Vec<true, true, true> op{pix, x, y, z, nside_v, base, hpgb_geom()};
            return hpgb_launch_map1(op, n, st, s);
        }
        OpPix2Vec<false, false, true> op{pix, x, y, z, nside_v, base, hpgb_geom()};
        return hpgb_launch_map1(op, n, st, s);
    }
    int c = hpgb_check_nside(nside, nest);
    if (c) return c;
    const hpgb_geom g = hpgb_make_geom(nside, nest);
    const bool vec = hpgb_aligned16(pix) && hpgb_aligned16(x) && hpgb_aligned16(y) && hpgb_aligned16(z);
    if (nest) {
        OpPix2Vec<true, true, false> op{pix, x, y, z, nullptr, base, g};
        return hpgb_launch_map(op, n, vec, st, s);
    }
    if (g.order >= 0) {
        OpPix2Vec<false, true, false> op{pix, x, y, z, nullptr, base, g};
        return hpgb_launch_map(op, n, vec, st, s);
    }
    OpPix2Vec<false, false, false> op{pix, x, y, z, nullptr, base, g};
    return hpgb_launch_map(op, n, vec, st, s);
}

// -------------------------------------------------------------------------------------------
// neighbors   (hpgeom/hpgeom.c:2928-2959)   8 B in + 64 B out per pixel, rows of 8 (row-major)
// One thread per pixel; the row leaves as four 128-bit streaming stores.
// -------------------------------------------------------------------------------------------
template <bool NEST, bool POW2, bool VAR>
struct OpNeighbors {
    const int64_t *pix;
    int64_t *out;
    const int64_t *nside_v;
    int64_t base;
    hpgb_geom g;
    typedef hpgb_no_ctx Ctx;
    typedef longlong2 In2;
    __device__ __forceinline__ Ctx prologue() const { return Ctx(); }
    __device__ __forceinline__ In2 load2(int64_t i) const { return hpgb_ld2(pix + i); }
    __device__ __forceinline__ void row(const hpgb_geom &gg, int64_t p, int64_t i, hpgb_bad_t &bad) const {
        int64_t r[8];
        if (!hpgb_pixel_ok(gg, p)) {
            bad.flag(base + i);
#pragma unroll
            for (int m = 0; m < 8; m++) r[m] = -1;
        } else {
            hpgb_neighbors<NEST, POW2>(gg, p, r);
        }
        int64_t *o = out + 8 * i;  // 64-byte rows: always 16-byte aligned when `out` is
#pragma unroll
        for (int m = 0; m < 8; m += 2) hpgb_st2(o + m, r[m], r[m + 1]);
    }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &, hpgb_bad_t &bad) const {
        row(g, v.x, i, bad);
        row(g, v.y, i + 1, bad);
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &, hpgb_bad_t &bad) const {
        hpgb_geom gg = g;
        if (VAR && !var_geom(nside_v, i, NEST, gg)) {
            bad.flag(base + i);
            for (int m = 0; m < 8; m++) out[8 * i + m] = -1;
            return;
        }
        row(gg, hpgb_ld1(pix + i), i, bad);
    }
};

int launch_neighbors(const int64_t *pix, int64_t *out, int64_t n, int64_t nside, const int64_t *nside_v, int nest,
                     int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!pix || !out || !st))) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;
    if (!hpgb_aligned16(out)) return HPGB_E_BAD_ARGUMENT;  // rows are written with 128-bit stores
    if (nside_v) {
        if (nest) {
            OpNeighbors<true, true, true> op{pix, out, nside_v, base, hpgb_geom()};
            return hpgb_launch_map1(op, n, st, s);
        }
        OpNeighbors<false, false, true> op{pix, out, nside_v, base, hpgb_geom()};
        return hpgb_launch_map1(op, n, st, s);
    }
    int c = hpgb_check_nside(nside, nest);
    if (c) return c;
    const hpgb_geom g = hpgb_make_geom(nside, nest);
    const bool vec = hpgb_aligned16(pix);
    if (nest) {
        OpNeighbors<true, true, false> op{pix, out, nullptr, base, g};
        return hpgb_launch_map(op, n, vec, st, s);
    }
    if (g.order >= 0) {
        OpNeighbors<false, true, false> op{pix, out, nullptr, base, g};
        return hpgb_launch_map(op, n, vec, st, s);
    }
    OpNeighbors<false, false, false> op{pix, out, nullptr, base, g};
    return hpgb_launch_map(op, n, vec, st, s);
}

// -------------------------------------------------------------------------------------------
// get_interpolation_weights   (hpgeom/hpgeom.c:3491-3523)   16 B in + 64 B out per point
// -------------------------------------------------------------------------------------------
template <bool NEST, bool POW2, bool LONLAT, bool DEGREES, bool VAR>
struct OpInterp {
    const double *a, *b;
    int64_t *pix;
    double *wgt;
    const int64_t *nside_v;
    int64_t base;
    hpgb_geom g;
    struct In2 {
        double2 a, b;
    };
    typedef const double *Ctx;
    __device__ __forceinline__ Ctx prologue() const { return hpgb_stage_table(); }
    __device__ __forceinline__ In2 load2(int64_t i) const { return In2{hpgb_ld2(a + i), hpgb_ld2(b + i)}; }
    __device__ __forceinline__ void row(const hpgb_geom &gg, double va, double vb, int64_t i, Ctx T, hpgb_bad_t &bad) const {
        int64_t p[4] = {-1, -1, -1, -1};
        double w[4] = {0.0, 0.0, 0.0, 0.0};
        double theta, phi;
        if (!hpgb_angles_exact<LONLAT, DEGREES>(va, vb, theta, phi))
            bad.flag(base + i);
        else
            hpgb_interpol<NEST, POW2>(gg, T, theta, phi, p, w);
        hpgb_st2(pix + 4 * i, p[0], p[1]);
        hpgb_st2(pix + 4 * i + 2, p[2], p[3]);
        hpgb_st2(wgt + 4 * i, w[0], w[1]);
        hpgb_st2(wgt + 4 * i + 2, w[2], w[3]);
    }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &T, hpgb_bad_t &bad) const {
        row(g, v.a.x, v.b.x, i, T, bad);
        row(g, v.a.y, v.b.y, i + 1, T, bad);
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &T, hpgb_bad_t &bad) const {
        hpgb_geom gg = g;
        if (VAR && !var_geom(nside_v, i, NEST, gg)) {
            bad.flag(base + i);
            for (int m = 0; m < 4; m++) {
                pix[4 * i + m] = -1;
                wgt[4 * i + m] = 0.0;
            }
            return;
        }
        row(gg, hpgb_ld1(a + i), hpgb_ld1(b + i), i, T, bad);
    }
};

template <bool NEST, bool LONLAT, bool DEGREES>
int launch_interp_t(const double *a, const double *b, int64_t *pix, double *wgt, int64_t n, int64_t nside,
                    const int64_t *nside_v, int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (nside_v) {
        OpInterp<NEST, NEST, LONLAT, DEGREES, true> op{a, b, pix, wgt, nside_v, base, hpgb_geom()};
        return hpgb_launch_map1(op, n, st, s);
    }
    const hpgb_geom g = hpgb_make_geom(nside, NEST);
    const bool vec = hpgb_aligned16(a) && hpgb_aligned16(b);
    if (NEST || g.order >= 0) {
        OpInterp<NEST, true, LONLAT, DEGREES, false> op{a, b, pix, wgt, nullptr, base, g};
        return hpgb_launch_map(op, n, vec, st, s);
    }
    OpInterp<false, false, LONLAT, DEGREES, false> op{a, b, pix, wgt, nullptr, base, g};
    return hpgb_launch_map(op, n, vec, st, s);
}

int launch_interp(const double *a, const double *b, int64_t *pix, double *wgt, int64_t n, int64_t nside,
                  const int64_t *nside_v, int nest, int lonlat, int degrees, int64_t base, hpgb_status_t *st,
                  cudaStream_t s) {
    if (n < 0 || (n > 0 && (!a || !b || !pix || !wgt || !st))) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;
    if (!hpgb_aligned16(pix) || !hpgb_aligned16(wgt)) return HPGB_E_BAD_ARGUMENT;
    if (!nside_v) {
        int c = hpgb_check_nside(nside, nest);
        if (c) return c;
    }
#define GO(N, L, D) return launch_interp_t<N, L, D>(a, b, pix, wgt, n, nside, nside_v, base, st, s)
    if (nest) {
        if (lonlat) {
            if (degrees) GO(true, true, true);
            GO(true, true, false);
        }
        GO(true, false, false);
    }
    if (lonlat) {
        if (degrees) GO(false, true, true);
        GO(false, true, false);
    }
    GO(false, false, false);
#undef GO
}

// -------------------------------------------------------------------------------------------
// upgrade_pixels   (hpgeom/hpgeom.py:515-556 fused: [ring->nest] -> children -> [nest->ring])
// One thread per OUTPUT pixel (coalesced 8 B..16 B stores); the parent is re-read through L1/L2
// by its 4^k children.  Element error classes (priority via the high bits of the atomicMin key):
//   class 0: pixel out of range              (the reference's ring_to_nest check)
//   class 1: nest pixel == npix-1            (the reference range-checks p+1, hpgeom.py:505)
// -------------------------------------------------------------------------------------------
#define HPGB_UPGRADE_CLASS1 ((int64_t)1 << 62)

template <bool NEST>
__global__ void __launch_bounds__(HPGB_BLOCK) k_upgrade(const int64_t *__restrict__ pix, int64_t *__restrict__ out,
                                                        const int64_t n_out, const int shift, const hpgb_geom g,
                                                        const hpgb_geom gu, const int64_t base, hpgb_status_t *st) {
    const int64_t stride = (int64_t)gridDim.x * HPGB_BLOCK;
    hpgb_bad_t bad;
    for (int64_t o = (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x; o < n_out; o += stride) {
        const int64_t i = o >> shift;
        const int64_t j = o & (((int64_t)1 << shift) - 1);
        int64_t p = __ldg(pix + i);
        int64_t r = -1;
        if (!hpgb_pixel_ok(g, p)) {
            bad.flag(base + i);
        } else {
            if (!NEST) p = hpgb_ring2nest(g, p);
            if (p == g.npix - 1) bad.flag(HPGB_UPGRADE_CLASS1 | (base + i));
            r = (p << shift) + j;
            if (!NEST) r = hpgb_nest2ring(gu, r);
        }
        __stcs(out + o, r);
    }
    bad.commit(st);
}

int launch_upgrade(const int64_t *pix, int64_t *out, int64_t n, int64_t nside, int64_t nside_upgrade, int nest,
                   int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!pix || !out || !st))) return HPGB_E_BAD_ARGUMENT;
    int c = hpgb_check_nside(nside, 1);
    if (c) return c;
    c = hpgb_check_nside(nside_upgrade, 1);
    if (c || nside_upgrade < nside) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;
    const hpgb_geom g = hpgb_make_geom(nside, 1), gu = hpgb_make_geom(nside_upgrade, 1);
    const int shift = 2 * (gu.order - g.order);
    const int64_t n_out = n << shift;
    static int occ[2] = {0, 0};
    if (!occ[nest ? 1 : 0])
        occ[nest ? 1 : 0] = nest ? hpgb_resident_ctas(k_upgrade<true>) : hpgb_resident_ctas(k_upgrade<false>);
    int64_t grid = (int64_t)hpgb_sm_count() * occ[nest ? 1 : 0];
    const int64_t want = (n_out + HPGB_BLOCK - 1) / HPGB_BLOCK;
    if (want < grid) grid = want;
    if (nest)
        k_upgrade<true><<<(unsigned)grid, HPGB_BLOCK, 0, s>>>(pix, out, n_out, shift, g, gu, base, st);
    else
        k_upgrade<false><<<(unsigned)grid, HPGB_BLOCK, 0, s>>>(pix, out, n_out, shift, g, gu, base, st);
    g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

// -------------------------------------------------------------------------------------------
// lonlat_to_thetaphi   (hpgeom/hpgeom.py:60-88, numpy formulas)   16 B in + 16 B out
// -------------------------------------------------------------------------------------------
template <bool DEGREES>
struct OpLonLat {
    const double *lon, *lat;
    double *theta, *phi;
    int64_t base;
    struct In2 {
        double2 lon, lat;
    };
    typedef hpgb_no_ctx Ctx;
    __device__ __forceinline__ Ctx prologue() const { return Ctx(); }
    __device__ __forceinline__ In2 load2(int64_t i) const { return In2{hpgb_ld2(lon + i), hpgb_ld2(lat + i)}; }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &, hpgb_bad_t &bad) const {
        double t0, p0, t1, p1;
        if (!hpgb_lonlat_to_thetaphi_np<DEGREES>(v.lon.x, v.lat.x, t0, p0)) bad.flag(base + i);
        if (!hpgb_lonlat_to_thetaphi_np<DEGREES>(v.lon.y, v.lat.y, t1, p1)) bad.flag(base + i + 1);
        hpgb_st2(theta + i, t0, t1);
        hpgb_st2(phi + i, p0, p1);
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &, hpgb_bad_t &bad) const {
        double t, p;
        if (!hpgb_lonlat_to_thetaphi_np<DEGREES>(hpgb_ld1(lon + i), hpgb_ld1(lat + i), t, p)) bad.flag(base + i);
        hpgb_st1(theta + i, t);
        hpgb_st1(phi + i, p);
    }
};

int launch_lonlat(const double *lon, const double *lat, double *theta, double *phi, int64_t n, int degrees,
                  int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!lon || !lat || !theta || !phi || !st))) return HPGB_E_BAD_ARGUMENT;
    const bool vec = hpgb_aligned16(lon) && hpgb_aligned16(lat) && hpgb_aligned16(theta) && hpgb_aligned16(phi);
    if (degrees) {
        OpLonLat<true> op{lon, lat, theta, phi, base};
        return hpgb_launch_map(op, n, vec, st, s);
    }
    OpLonLat<false> op{lon, lat, theta, phi, base};
    return hpgb_launch_map(op, n, vec, st, s);
}

typedef const char *cc;

}  // namespace

// =============================================================================================
// C ABI (include/hpgb.h)
// =============================================================================================
extern "C" {

#define S(x) ((cudaStream_t)(x))

int hpgb_nest_to_ring(const int64_t *d_nest, int64_t *d_ring, int64_t n, int64_t nside, const int64_t *d_nside_v,
                      hpgb_status_t *d_status, void *stream) {
    return launch_reorder<true>(d_nest, d_ring, n, nside, d_nside_v, 0, d_status, S(stream));
}
int hpgb_ring_to_nest(const int64_t *d_ring, int64_t *d_nest, int64_t n, int64_t nside, const int64_t *d_nside_v,
                      hpgb_status_t *d_status, void *stream) {
    return launch_reorder<false>(d_ring, d_nest, n, nside, d_nside_v, 0, d_status, S(stream));
}
int hpgb_angle_to_pixel(const double *d_a, const double *d_b, int64_t *d_pix, int64_t n, int64_t nside,
                        const int64_t *d_nside_v, int nest, int lonlat, int degrees, hpgb_status_t *d_status,
                        void *stream) {
    return launch_a2p(d_a, d_b, d_pix, n, nside, d_nside_v, nest, lonlat, degrees, 0, d_status, S(stream));
}
int hpgb_pixel_to_angle(const int64_t *d_pix, double *d_a, double *d_b, int64_t n, int64_t nside,
                        const int64_t *d_nside_v, int nest, int lonlat, int degrees, hpgb_status_t *d_status,
                        void *stream) {
    return launch_p2a(d_pix, d_a, d_b, n, nside, d_nside_v, nest, lonlat, degrees, 0, d_status, S(stream));
}
int hpgb_vector_to_pixel(const double *d_x, const double *d_y, const double *d_z, int64_t *d_pix, int64_t n,
                         int64_t nside, const int64_t *d_nside_v, int nest, hpgb_status_t *d_status, void *stream) {
    return launch_v2p(d_x, d_y, d_z, d_pix, n, nside, d_nside_v, nest, 0, d_status, S(stream));
}
int hpgb_pixel_to_vector(const int64_t *d_pix, double *d_x, double *d_y, double *d_z, int64_t n, int64_t nside,
                         const int64_t *d_nside_v, int nest, hpgb_status_t *d_status, void *stream) {
    return launch_p2v(d_pix, d_x, d_y, d_z, n, nside, d_nside_v, nest, 0, d_status, S(stream));
}
int hpgb_neighbors(const int64_t *d_pix, int64_t *d_out_n8, int64_t n, int64_t nside, const int64_t *d_nside_v,
                   int nest, hpgb_status_t *d_status, void *stream) {
    return launch_neighbors(d_pix, d_out_n8, n, nside, d_nside_v, nest, 0, d_status, S(stream));
}
int hpgb_interp_weights(const double *d_a, const double *d_b, int64_t *d_pix_n4, double *d_wgt_n4, int64_t n,
                        int64_t nside, const int64_t *d_nside_v, int nest, int lonlat, int degrees,
                        hpgb_status_t *d_status, void *stream) {
    return launch_interp(d_a, d_b, d_pix_n4, d_wgt_n4, n, nside, d_nside_v, nest, lonlat, degrees, 0, d_status, S(stream));
}
int hpgb_upgrade_pixels(const int64_t *d_pix, int64_t *d_out, int64_t n, int64_t nside, int64_t nside_upgrade,
                        int nest, hpgb_status_t *d_status, void *stream) {
    return launch_upgrade(d_pix, d_out, n, nside, nside_upgrade, nest, 0, d_status, S(stream));
}
int hpgb_lonlat_to_thetaphi(const double *d_lon, const double *d_lat, double *d_theta, double *d_phi, int64_t n,
                            int degrees, hpgb_status_t *d_status, void *stream) {
    return launch_lonlat(d_lon, d_lat, d_theta, d_phi, n, degrees, 0, d_status, S(stream));
}

// ---- host-pointer entry points: operands in C-argument order, then the optional nside array ----
#define IN(p, b) {(cc)(p), nullptr, (b)}
#define OUT(p, b) {nullptr, (char *)(p), (b)}
#define PIPE(arrs, narr, call)                                                                          \
    return hpgb_host_pipeline(device, n, arrs, narr,                                                    \
                              [&](void **d, int64_t base, int64_t cn, hpgb_status_t *st, cudaStream_t s) { return call; }, \
                              first_bad)
#define NSV(k) (nside_v ? (const int64_t *)d[k] : nullptr)

int hpgb_host_nest_to_ring(const int64_t *nest, int64_t *ring, int64_t n, int64_t nside, const int64_t *nside_v,
                           int device, int64_t *first_bad) {
    hpgb_pipe_arr arrs[3] = {IN(nest, 8), OUT(ring, 8), IN(nside_v, 8)};
    PIPE(arrs, nside_v ? 3 : 2, launch_reorder<true>((const int64_t *)d[0], (int64_t *)d[1], cn, nside, NSV(2), base, st, s));
}
int hpgb_host_ring_to_nest(const int64_t *ring, int64_t *nest, int64_t n, int64_t nside, const int64_t *nside_v,
                           int device, int64_t *first_bad) {
    hpgb_pipe_arr arrs[3] = {IN(ring, 8), OUT(nest, 8), IN(nside_v, 8)};
    PIPE(arrs, nside_v ? 3 : 2, launch_reorder<false>((const int64_t *)d[0], (int64_t *)d[1], cn, nside, NSV(2), base, st, s));
}
int hpgb_host_angle_to_pixel(const double *a, const double *b, int64_t *pix, int64_t n, int64_t nside,
                             const int64_t *nside_v, int nest, int lonlat, int degrees, int device, int64_t *first_bad) {
    hpgb_pipe_arr arrs[4] = {IN(a, 8), IN(b, 8), OUT(pix, 8), IN(nside_v, 8)};
    PIPE(arrs, nside_v ? 4 : 3,
         launch_a2p((const double *)d[0], (const double *)d[1], (int64_t *)d[2], cn, nside, NSV(3), nest, lonlat, degrees, base, st, s));
}
int hpgb_host_pixel_to_angle(const int64_t *pix, double *a, double *b, int64_t n, int64_t nside,
                             const int64_t *nside_v, int nest, int lonlat, int degrees, int device, int64_t *first_bad) {
    hpgb_pipe_arr arrs[4] = {IN(pix, 8), OUT(a, 8), OUT(b, 8), IN(nside_v, 8)};
    PIPE(arrs, nside_v ? 4 : 3,
         launch_p2a((const int64_t *)d[0], (double *)d[1], (double *)d[2], cn, nside, NSV(3), nest, lonlat, degrees, base, st, s));
}
int hpgb_host_vector_to_pixel(const double *x, const double *y, const double *z, int64_t *pix, int64_t n,
                              int64_t nside, const int64_t *nside_v, int nest, int device, int64_t *first_bad) {
    hpgb_pipe_arr arrs[5] = {IN(x, 8), IN(y, 8), IN(z, 8), OUT(pix, 8), IN(nside_v, 8)};
    PIPE(arrs, nside_v ? 5 : 4,
         launch_v2p((const double *)d[0], (const double *)d[1], (const double *)d[2], (int64_t *)d[3], cn, nside, NSV(4), nest, base, st, s));
}
int hpgb_host_pixel_to_vector(const int64_t *pix, double *x, double *y, double *z, int64_t n, int64_t nside,
                              const int64_t *nside_v, int nest, int device, int64_t *first_bad) {
    hpgb_pipe_arr arrs[5] = {IN(pix, 8), OUT(x, 8), OUT(y, 8), OUT(z, 8), IN(nside_v, 8)};
    PIPE(arrs, nside_v ? 5 : 4,
         launch_p2v((const int64_t *)d[0], (double *)d[1], (double *)d[2], (double *)d[3], cn, nside, NSV(4), nest, base, st, s));
}
int hpgb_host_neighbors(const int64_t *pix, int64_t *out_n8, int64_t n, int64_t nside, const int64_t *nside_v,
                        int nest, int device, int64_t *first_bad) {
    hpgb_pipe_arr arrs[3] = {IN(pix, 8), OUT(out_n8, 64), IN(nside_v, 8)};
    PIPE(arrs, nside_v ? 3 : 2, launch_neighbors((const int64_t *)d[0], (int64_t *)d[1], cn, nside, NSV(2), nest, base, st, s));
}
int hpgb_host_interp_weights(const double *a, const double *b, int64_t *pix_n4, double *wgt_n4, int64_t n,
                             int64_t nside, const int64_t *nside_v, int nest, int lonlat, int degrees, int device,
                             int64_t *first_bad) {
    hpgb_pipe_arr arrs[5] = {IN(a, 8), IN(b, 8), OUT(pix_n4, 32), OUT(wgt_n4, 32), IN(nside_v, 8)};
    PIPE(arrs, nside_v ? 5 : 4,
         launch_interp((const double *)d[0], (const double *)d[1], (int64_t *)d[2], (double *)d[3], cn, nside, NSV(4), nest, lonlat, degrees, base, st, s));
}
int hpgb_host_upgrade_pixels(const int64_t *pix, int64_t *out, int64_t n, int64_t nside, int64_t nside_upgrade,
                             int nest, int device, int64_t *first_bad) {
    if (hpgb_check_nside(nside, 1)) return hpgb_check_nside(nside, 1);
    if (hpgb_check_nside(nside_upgrade, 1) || nside_upgrade < nside) return HPGB_E_BAD_ARGUMENT;
    const int64_t ratio = (nside_upgrade / nside) * (nside_upgrade / nside);
    hpgb_pipe_arr arrs[2] = {IN(pix, 8), OUT(out, 8 * ratio)};
    PIPE(arrs, 2, launch_upgrade((const int64_t *)d[0], (int64_t *)d[1], cn, nside, nside_upgrade, nest, base, st, s));
}
int hpgb_host_lonlat_to_thetaphi(const double *lon, const double *lat, double *theta, double *phi, int64_t n,
                                 int degrees, int device, int64_t *first_bad) {
    hpgb_pipe_arr arrs[4] = {IN(lon, 8), IN(lat, 8), OUT(theta, 8), OUT(phi, 8)};
    PIPE(arrs, 4, launch_lonlat((const double *)d[0], (const double *)d[1], (double *)d[2], (double *)d[3], cn, degrees, base, st, s));
}

}  // extern "C"
