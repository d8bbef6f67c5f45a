// hpgb_kernels.cu -- sm_100a kernels + C-ABI entry points (include/hpgb.h) for hpgeom's
// elementwise HEALPix conversion path.  One Op struct per API function (see hpgb_launch.cuh
// for the kernel skeleton they plug into); templated on the flags that the reference tests per
// element (nest / lonlat / degrees, hpgeom/hpgeom.c:143-153) so those branches are
// compile-time here.
#include <cuda_runtime.h>
#include <stdint.h>

#include "hpgb_core.h"
#include "hpgb_host.h"
#include "hpgb_launch.cuh"

namespace {

// -------------------------------------------------------------------------------------------
// nest_to_ring / ring_to_nest   (reference loops: hpgeom/hpgeom.c:1429-1450, :1699-1720)
// 8 B in + 8 B out per pixel, integer only.
// -------------------------------------------------------------------------------------------
template <bool TO_RING, bool VARNSIDE>
struct OpReorder {
    const int64_t *in;
    int64_t *out;
    const int64_t *nside_v;
    int64_t base;  // global index of element 0 (chunked host pipeline)
    hpgb_geom g;
    typedef longlong2 In2;

    __device__ __forceinline__ void prologue() const {}
    __device__ __forceinline__ int64_t conv(const hpgb_geom &gg, int64_t p, int64_t i, hpgb_bad_t &bad) const {
        if (!hpgb_pixel_ok(gg, p)) {
            bad.flag(base + i);
            return -1;
        }
        return TO_RING ? hpgb_nest2ring(gg, p) : hpgb_ring2nest(gg, p);
    }
    __device__ __forceinline__ In2 load2(int64_t i) const { return hpgb_ld2(in + i); }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, hpgb_bad_t &bad) const {
        hpgb_st2(out + i, conv(g, v.x, i, bad), conv(g, v.y, i + 1, bad));
    }
    __device__ __forceinline__ void run1(int64_t i, hpgb_bad_t &bad) const {
        if (VARNSIDE) {
            const int64_t ns = hpgb_ld1(nside_v + i);
            if (hpgb_check_nside(ns, 1)) {
                bad.flag(base + i);
                hpgb_st1(out + i, -1);
                return;
            }
            hpgb_st1(out + i, conv(hpgb_make_geom(ns, 1), hpgb_ld1(in + i), i, bad));
        } else {
            hpgb_st1(out + i, conv(g, hpgb_ld1(in + i), i, bad));
        }
    }
};

template <bool TO_RING>
int launch_reorder(const int64_t *in, int64_t *out, int64_t n, int64_t nside, const int64_t *nside_v,
                   int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!in || !out || !st))) return HPGB_E_BAD_ARGUMENT;
    if (nside_v) {
        OpReorder<TO_RING, true> op{in, out, nside_v, base, hpgb_geom()};
        return hpgb_launch_map(op, n, false, st, s);
    }
    if (n == 0) return HPGB_OK;  // the reference checks nside inside the loop: no elements, no error
    int c = hpgb_check_nside(nside, 1);
    if (c) return c;
    OpReorder<TO_RING, false> op{in, out, nullptr, base, hpgb_make_geom(nside, 1)};
    return hpgb_launch_map(op, n, hpgb_aligned16(in) && hpgb_aligned16(out), st, s);
}

template <bool TO_RING>
int host_reorder(const int64_t *in, int64_t *out, int64_t n, int64_t nside, const int64_t *nside_v,
                 int device, int64_t *first_bad) {
    hpgb_pipe_arr arrs[3] = {{(const char *)in, nullptr, 8}, {nullptr, (char *)out, 8}, {(const char *)nside_v, nullptr, 8}};
    return hpgb_host_pipeline(device, n, arrs, nside_v ? 3 : 2,
                              [&](void **d, int64_t base, int64_t cn, hpgb_status_t *st, cudaStream_t s) {
                                  return launch_reorder<TO_RING>((const int64_t *)d[0], (int64_t *)d[1], cn, nside,
                                                                 nside_v ? (const int64_t *)d[2] : nullptr, base, st, s);
                              },
                              first_bad);
}

}  // namespace

extern "C" {

int hpgb_nest_to_ring(const int64_t *d_nest, int64_t *d_ring, int64_t n, int64_t nside, const int64_t *d_nside_v,
                      hpgb_status_t *d_status, void *stream) {
    return launch_reorder<true>(d_nest, d_ring, n, nside, d_nside_v, 0, d_status, (cudaStream_t)stream);
}
int hpgb_ring_to_nest(const int64_t *d_ring, int64_t *d_nest, int64_t n, int64_t nside, const int64_t *d_nside_v,
                      hpgb_status_t *d_status, void *stream) {
    return launch_reorder<false>(d_ring, d_nest, n, nside, d_nside_v, 0, d_status, (cudaStream_t)stream);
}
int hpgb_host_nest_to_ring(const int64_t *nest, int64_t *ring, int64_t n, int64_t nside, const int64_t *nside_v,
                           int device, int64_t *first_bad) {
    return host_reorder<true>(nest, ring, n, nside, nside_v, device, first_bad);
}
int hpgb_host_ring_to_nest(const int64_t *ring, int64_t *nest, int64_t n, int64_t nside, const int64_t *nside_v,
                           int device, int64_t *first_bad) {
    return host_reorder<false>(ring, nest, n, nside, nside_v, device, first_bad);
}

}  // extern "C"
