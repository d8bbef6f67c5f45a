// k_helpers.cu -- the small elementwise helpers that sit next to the conversion path in the reference's Python
// module (hpgeom/hpgeom.py:91-112 thetaphi_to_lonlat, :364-393 angle_to_vector, :396-422 vector_to_angle) and the
// full-map RING <-> NEST permutation (:425-462 reorder).  The reference evaluates them with numpy on the host; here
// each is one kernel, so that numpy callers do not compute on the CPU and CUDA-tensor callers do not fall back on
// eager torch ops (whose sin / cos / acos / atan2 are not the reference's).
//
// Parity: thetaphi_to_lonlat is pure IEEE (one subtraction, one multiplication by the constant 180/pi) and bit-exact.
// angle_to_vector uses numpy's sin / cos = glibc's, correctly rounded for ~99.9 % of arguments: hpgb_sincos_sd
// (<= 0.51 ulp) returns the same bits wherever glibc rounds correctly, the three products are IEEE.  vector_to_angle
// uses numpy's arccos / arctan2, which on AVX-512 hosts are SVML routines (<= 1 ulp, NOT correctly rounded): there is
// no bit-exact target, the correctly rounded acos / atan2 of hpgb_math.h stay within 1 ulp of it (tests: <= 2 ulp).
#include <cuda_runtime.h>
#include <stdint.h>

#include "hpgb_core.h"
#include "hpgb_fp.h"
#include "hpgb_host.h"
#include "hpgb_internal.h"
#include "hpgb_launch.cuh"
#include "hpgb_reorder.h"

namespace {

// error classes of angle_to_vector's theta / phi checks (hpgeom/hpgeom.py:342-361: np.any over ALL thetas first,
// then over all phis): the class rides in the top bits of the atomicMin key, so a bad theta anywhere wins
#define HPGB_A2V_CLASS_PHI ((int64_t)1 << 62)

// -------------------------------------------------------------------------------------------
// thetaphi_to_lonlat   16 B in + 16 B out
// -------------------------------------------------------------------------------------------
template <bool DEGREES>
struct OpThetaPhi2LonLat {
    const double *theta, *phi;
    double *lon, *lat;
    struct In2 {
        double2 a, b;
    };
    typedef hpgb_no_ctx Ctx;
    HPGB_DEFAULT_RUN4
    HPGB_STREAM_IN_F64x2(theta, phi, 4)
    __device__ __forceinline__ Ctx prologue() const { return Ctx(); }
    __device__ __forceinline__ In2 load2(int64_t i) const { return In2{hpgb_ld2(theta + i), hpgb_ld2(phi + i)}; }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &, hpgb_bad_t &) const {
        double lo0, la0, lo1, la1;
        hpgb_thetaphi_out<true, DEGREES>(v.a.x, v.b.x, lo0, la0);
        hpgb_thetaphi_out<true, DEGREES>(v.a.y, v.b.y, lo1, la1);
        hpgb_st2(lon + i, lo0, lo1);
        hpgb_st2(lat + i, la0, la1);
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &, hpgb_bad_t &) const {
        double lo, la;
        hpgb_thetaphi_out<true, DEGREES>(hpgb_ld1(theta + i), hpgb_ld1(phi + i), lo, la);
        hpgb_st1(lon + i, lo);
        hpgb_st1(lat + i, la);
    }
};

// -------------------------------------------------------------------------------------------
// angle_to_vector   16 B in + 24 B out, rows of 3 (row-major)
// -------------------------------------------------------------------------------------------
template <bool LONLAT, bool DEGREES>
struct OpAng2Vec {
    const double *a, *b;
    double *vec;
    int64_t base;
    struct In2 {
        double2 a, b;
    };
    typedef const double *Ctx;
    HPGB_DEFAULT_RUN4
    HPGB_STREAM_IN_F64x2(a, b, 2)
    __device__ __forceinline__ Ctx prologue() const { return hpgb_stage_table(); }
    __device__ __forceinline__ In2 load2(int64_t i) const { return In2{hpgb_ld2(a + i), hpgb_ld2(b + i)}; }
    __device__ __forceinline__ void one(double va, double vb, int64_t i, Ctx T, hpgb_bad_t &bad, double &x, double &y, double &z) const {
        double theta, phi;
        if (LONLAT) {
            if (!hpgb_lonlat_to_thetaphi_np<DEGREES>(va, vb, theta, phi)) bad.flag(base + i);
        } else {
            theta = va, phi = vb;
            if (va < 0.0 || va > HPGB_PI)
                bad.flag(base + i);
            else if (vb < 0.0 || vb > 2 * HPGB_PI)
                bad.flag(HPGB_A2V_CLASS_PHI | (base + i));
        }
        // out-of-range (or non-finite) angles still get defined arithmetic: clamp the argument of the table
        // routine, whose reduction covers |x| < 2^20
        const double th = fabs(theta) < 1e5 ? theta : 0.0, ph = fabs(phi) < 1e5 ? phi : 0.0;
        double st, ct, sp, cp;
        hpgb_sincos_sd(T, th, st, ct);
        hpgb_sincos_sd(T, ph, sp, cp);
        x = st * cp;
        y = st * sp;
        z = ct;
    }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &T, hpgb_bad_t &bad) const {
        double x0, y0, z0, x1, y1, z1;
        one(v.a.x, v.b.x, i, T, bad, x0, y0, z0);
        one(v.a.y, v.b.y, i + 1, T, bad, x1, y1, z1);
        double *o = vec + 3 * i;  // i even: 48-byte aligned when vec is 16-byte aligned
        hpgb_st2(o, x0, y0);
        hpgb_st2(o + 2, z0, x1);
        hpgb_st2(o + 4, y1, z1);
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &T, hpgb_bad_t &bad) const {
        double x, y, z;
        one(hpgb_ld1(a + i), hpgb_ld1(b + i), i, T, bad, x, y, z);
        hpgb_st1(vec + 3 * i, x);
        hpgb_st1(vec + 3 * i + 1, y);
        hpgb_st1(vec + 3 * i + 2, z);
    }
};

// -------------------------------------------------------------------------------------------
// vector_to_angle   24 B in (rows of 3) + 16 B out.  No validation (as the reference).
// -------------------------------------------------------------------------------------------
// acos over the whole domain: the double-double routine for |z| <= 0.99, beyond it atan2(sqrt((1 - z)(1 + z)), z)
// with (1 - |z|) exact and the product kept as a double-double under the square root
__device__ __forceinline__ double acos_full(const double *T, double z) {
    const double az = fabs(z);
    if (!(az < 1.0)) return z > 0.0 ? 0.0 : (z < 0.0 ? HPGB_PI : z);  // +-1 (or NaN -> NaN; |z| > 1 cannot come from z / |v|)
    if (az <= 0.99) return hpgb_acos_cr(T, z);
    const double u = 1.0 - az;  // exact
    double w, we, p, pe;
    hpgb_two_sum(1.0, az, w, we);
    hpgb_two_prod(u, w, p, pe);
    const double s = sqrt(p);
    const double sc = s + (fma(u, we, pe) + fma(-s, s, p)) / (2.0 * s);  // one Newton correction from the exact residual
    return hpgb_atan2_cr(T, sc, z);
}

template <bool LONLAT, bool DEGREES>
struct OpVec2Ang {
    const double *vec;
    double *a, *b;
    struct In2 {
        double2 p, q, r;  // x0 y0 | z0 x1 | y1 z1
    };
    typedef const double *Ctx;
    HPGB_DEFAULT_RUN4
    __device__ __forceinline__ Ctx prologue() const { return hpgb_stage_table(); }
    __device__ __forceinline__ In2 load2(int64_t i) const {
        const double *s = vec + 3 * i;
        return In2{hpgb_ld2(s), hpgb_ld2(s + 2), hpgb_ld2(s + 4)};
    }
    __device__ __forceinline__ void one(double x, double y, double z, Ctx T, double &oa, double &ob) const {
        const double norm = sqrt((x * x + y * y) + z * z);  // np.sum(np.square(vec), axis=1): left to right
        const double theta = acos_full(T, z / norm);
        double phi = ((x == 0.0) && (y == 0.0)) ? atan2(y, x) : hpgb_atan2_any(T, y, x);
        // numpy's float remainder by 2 pi: the result takes the divisor's sign (phi is already in (-pi, pi])
        if (phi < 0.0) {
            phi = phi + 2 * HPGB_PI;
        } else if (phi == 0.0) {
            phi = 0.0;  // -0.0 % m = +0.0
        }
        hpgb_thetaphi_out<LONLAT, DEGREES>(theta, phi, oa, ob);
    }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &T, hpgb_bad_t &) const {
        double a0, b0, a1, b1;
        one(v.p.x, v.p.y, v.q.x, T, a0, b0);
        one(v.q.y, v.r.x, v.r.y, T, a1, b1);
        hpgb_st2(a + i, a0, a1);
        hpgb_st2(b + i, b0, b1);
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &T, hpgb_bad_t &) const {
        double oa, ob;
        one(hpgb_ld1(vec + 3 * i), hpgb_ld1(vec + 3 * i + 1), hpgb_ld1(vec + 3 * i + 2), T, oa, ob);
        hpgb_st1(a + i, oa);
        hpgb_st1(b + i, ob);
    }
};

// -------------------------------------------------------------------------------------------
// reorder: out[ring_to_nest(i)] = in[i] (or the converse) for a FULL map of 12 nside^2 elements of 1 / 2 / 4 / 8
// bytes.  One pass, 2 x element size of traffic per pixel:
// a CTA owns a B x B block of (ix, iy) of one base face (B = min(nside, 64)).  In NEST order that block is ONE
// contiguous run of B^2 elements (Morton order); in RING order it is 2B - 1 runs, one per ring the block touches
// (the pixels with ix + iy = const, consecutive along their ring).  The block goes through shared memory, indexed
// by the pixel's position in the NEST run: the NEST side moves as one linear copy, the RING side as run-wise
// copies (a warp per run, lanes along it), each run located by ONE xyf2ring of its first pixel.
// -------------------------------------------------------------------------------------------
template <typename E, bool TO_NEST>
__global__ void __launch_bounds__(256) k_reorder_map(const E *__restrict__ in, E *__restrict__ out, const hpgb_geom g, const hpgb_rk c,
                                                     const int lb /* log2 B */) {
    extern __shared__ __align__(16) unsigned char hpgb_reorder_smem[];
    E *tile = reinterpret_cast<E *>(hpgb_reorder_smem);
    const uint32_t B = 1u << lb;
    const uint32_t bpf = (uint32_t)(g.nside >> lb);  // blocks per face edge
    const uint32_t face = blockIdx.x / (bpf * bpf), brem = blockIdx.x % (bpf * bpf);
    // blocks are numbered in Morton order inside the face so that consecutive CTAs write / read consecutive NEST runs
    uint32_t bx, by;
    hpgb_deinterleave((uint64_t)brem, bx, by);
    const uint32_t x0 = bx << lb, y0 = by << lb;
    const int64_t nest0 = hpgb_xyf2nest(g.order, x0, y0, face);  // first element of the block's NEST run
    const uint32_t nb = B * B;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    if (!TO_NEST) {  // NEST -> RING: linear read
        for (uint32_t k = threadIdx.x; k < nb; k += blockDim.x) tile[k] = in[nest0 + k];
        __syncthreads();
    }
    // the runs: diagonal d = dx + dy (0 .. 2B-2), ring = const - d; the run starts at the block's WEST-most pixel of
    // that diagonal, i.e. smallest dx: (dx, dy) = (max(0, d - (B-1)), min(d, B-1)), and moves (dx + 1, dy - 1) = east
    for (uint32_t d = warp; d < 2 * B - 1; d += nwarp) {
        const uint32_t dx0 = d < B ? 0u : d - (B - 1u), dy0 = d < B ? d : B - 1u;
        const uint32_t len = d < B ? d + 1u : 2u * B - 1u - d;
        const int64_t r0 = hpgb_xyf2ring_k(c, x0 + dx0, y0 + dy0, face);
        // a run never wraps around its ring inside one face EXCEPT in face 4..7's belt rows crossing phi = 0:
        // hpgb_xyf2ring_k returns the wrapped index of the first pixel only, so locate the wrap: pixels whose
        // in-ring position passes the ring's end continue at the ring's start (same ring: subtract its length)
        int64_t ring_start, ring_len;
        {
            // ring of this diagonal: jr = (face/4 + 2) nside - (x + y) - 1 (1-based), hpgeom/healpix_geom.c:394-408
            const int64_t ns = g.nside, jr = (int64_t)(2u + (face >> 2)) * ns - (int64_t)(x0 + y0 + d) - 1;
            if (jr < ns) {
                ring_len = 4 * jr;
                ring_start = 2 * jr * (jr - 1);
            } else if (jr > 3 * ns) {
                const int64_t nr = 4 * ns - jr;
                ring_len = 4 * nr;
                ring_start = g.npix - 2 * nr * (nr + 1);
            } else {
                ring_len = 4 * ns;
                ring_start = g.ncap + (jr - ns) * 4 * ns;
            }
        }
        for (uint32_t k = lane; k < len; k += 32) {
            int64_t r = r0 + k;
            if (r >= ring_start + ring_len) r -= ring_len;
            const uint32_t m = (uint32_t)hpgb_interleave(dx0 + k, dy0 - k);  // position in the NEST run
            if (TO_NEST)
                tile[m] = in[r];
            else
                out[r] = tile[m];
        }
    }
    if (TO_NEST) {
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < nb; k += blockDim.x) out[nest0 + k] = tile[k];
    }
}

template <typename E>
int launch_reorder_map_t(const void *in, void *out, const hpgb_geom &g, int to_nest, cudaStream_t s) {
    int lb = g.order < 6 ? g.order : 6;
    while (lb > 0 && ((size_t)sizeof(E) << (2 * lb)) > 48 * 1024) lb--;  // static opt-in free: <= 48 KiB
    const int64_t blocks = 12 * ((g.nside >> lb) * (g.nside >> lb));
    const size_t smem = (size_t)sizeof(E) << (2 * lb);
    const hpgb_rk c = hpgb_make_rk(g);
    if (to_nest)
        k_reorder_map<E, true><<<(unsigned)blocks, 256, smem, s>>>((const E *)in, (E *)out, g, c, lb);
    else
        k_reorder_map<E, false><<<(unsigned)blocks, 256, smem, s>>>((const E *)in, (E *)out, g, c, lb);
    g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

}  // namespace

int launch_thetaphi_to_lonlat(const double *theta, const double *phi, double *lon, double *lat, int64_t n, int degrees,
                              hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!theta || !phi || !lon || !lat || !st))) return HPGB_E_BAD_ARGUMENT;
    const bool vec = hpgb_aligned16(theta) && hpgb_aligned16(phi) && hpgb_aligned16(lon) && hpgb_aligned16(lat);
    if (degrees) {
        OpThetaPhi2LonLat<true> op{theta, phi, lon, lat};
        return hpgb_launch_auto(op, n, vec, st, s);
    }
    OpThetaPhi2LonLat<false> op{theta, phi, lon, lat};
    return hpgb_launch_auto(op, n, vec, st, s);
}

int launch_angle_to_vector(const double *a, const double *b, double *vec3, int64_t n, int lonlat, int degrees, int64_t base,
                           hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!a || !b || !vec3 || !st))) return HPGB_E_BAD_ARGUMENT;
    const bool vec = hpgb_aligned16(a) && hpgb_aligned16(b) && hpgb_aligned16(vec3);
#define GO(L, D)                                  \
    {                                             \
        OpAng2Vec<L, D> op{a, b, vec3, base};     \
        return hpgb_launch_auto(op, n, vec, st, s); \
    }
    if (lonlat) {
        if (degrees) GO(true, true) else GO(true, false)
    }
    GO(false, false)
#undef GO
}

int launch_vector_to_angle(const double *vec3, double *a, double *b, int64_t n, int lonlat, int degrees, hpgb_status_t *st,
                           cudaStream_t s) {
    if (n < 0 || (n > 0 && (!vec3 || !a || !b || !st))) return HPGB_E_BAD_ARGUMENT;
    const bool vec = hpgb_aligned16(vec3) && hpgb_aligned16(a) && hpgb_aligned16(b);
#define GO(L, D)                                \
    {                                           \
        OpVec2Ang<L, D> op{vec3, a, b};         \
        return hpgb_launch_map(op, n, vec, st, s); \
    }
    if (lonlat) {
        if (degrees) GO(true, true) else GO(true, false)
    }
    GO(false, false)
#undef GO
}

int launch_reorder_map(const void *in, void *out, int64_t npix, int elem_bytes, int ring_to_nest, cudaStream_t s) {
    if (npix <= 0 || !in || !out || in == out) return HPGB_E_BAD_ARGUMENT;
    // npix = 12 nside^2 with nside a power of two <= 2^29 (hpgeom/hpgeom.py:443-446: npixel_to_nside + nside_to_order)
    if (npix % 12) return HPGB_E_BAD_ARGUMENT;
    const int64_t np = npix / 12;
    const int64_t nside = (int64_t)llround(sqrt((double)np));
    if (nside * nside != np) return HPGB_E_BAD_ARGUMENT;
    const int c = hpgb_check_nside(nside, 1);
    if (c) return c;
    const hpgb_geom g = hpgb_make_geom(nside, 1);
    if (((uintptr_t)in | (uintptr_t)out) & (uintptr_t)(elem_bytes - 1)) return HPGB_E_BAD_ARGUMENT;
    switch (elem_bytes) {
        case 1: return launch_reorder_map_t<uint8_t>(in, out, g, ring_to_nest, s);
        case 2: return launch_reorder_map_t<uint16_t>(in, out, g, ring_to_nest, s);
        case 4: return launch_reorder_map_t<uint32_t>(in, out, g, ring_to_nest, s);
        case 8: return launch_reorder_map_t<uint64_t>(in, out, g, ring_to_nest, s);
        case 16: return launch_reorder_map_t<ulonglong2>(in, out, g, ring_to_nest, s);
        default: return HPGB_E_BAD_ARGUMENT;
    }
}
