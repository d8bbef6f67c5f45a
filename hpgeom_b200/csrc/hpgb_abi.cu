// hpgb_abi.cu -- the C ABI of include/hpgb.h: thin extern "C" wrappers over the kernel launchers
// (device-pointer entry points) and over the chunked host pipeline (host-pointer entry points).
#include <cuda_runtime.h>
#include <stdint.h>

#include "hpgb_core.h"
#include "hpgb_host.h"
#include "hpgb_internal.h"

typedef const char *cc;

// =============================================================================================
// C ABI (include/hpgb.h)
// =============================================================================================
extern "C" {

#define S(x) ((cudaStream_t)(x))

int hpgb_nest_to_ring(const int64_t *d_nest, int64_t *d_ring, int64_t n, int64_t nside, const int64_t *d_nside_v,
                      hpgb_status_t *d_status, void *stream) {
    return launch_n2r(d_nest, d_ring, n, nside, d_nside_v, 0, d_status, S(stream));
}
int hpgb_ring_to_nest(const int64_t *d_ring, int64_t *d_nest, int64_t n, int64_t nside, const int64_t *d_nside_v,
                      hpgb_status_t *d_status, void *stream) {
    return launch_r2n(d_ring, d_nest, n, nside, d_nside_v, 0, d_status, S(stream));
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

int hpgb_boundaries(const int64_t *d_pix, double *d_a, double *d_b, int64_t n, int64_t nside, const int64_t *d_nside_v,
                    int64_t step, int nest, int lonlat, int degrees, hpgb_status_t *d_status, void *stream) {
    return launch_boundaries(d_pix, d_a, d_b, n, nside, d_nside_v, step, nest, lonlat, degrees, 0, d_status, S(stream));
}
int hpgb_max_pixel_radius(const int64_t *d_nside, double *d_radius, int64_t n, int degrees, hpgb_status_t *d_status,
                          void *stream) {
    return launch_max_pixrad(d_nside, d_radius, n, degrees, 0, d_status, S(stream));
}

int hpgb_thetaphi_to_lonlat(const double *d_theta, const double *d_phi, double *d_lon, double *d_lat, int64_t n, int degrees,
                            hpgb_status_t *d_status, void *stream) {
    return launch_thetaphi_to_lonlat(d_theta, d_phi, d_lon, d_lat, n, degrees, d_status, S(stream));
}
int hpgb_angle_to_vector(const double *d_a, const double *d_b, double *d_vec_n3, int64_t n, int lonlat, int degrees,
                         hpgb_status_t *d_status, void *stream) {
    return launch_angle_to_vector(d_a, d_b, d_vec_n3, n, lonlat, degrees, 0, d_status, S(stream));
}
int hpgb_vector_to_angle(const double *d_vec_n3, double *d_a, double *d_b, int64_t n, int lonlat, int degrees,
                         hpgb_status_t *d_status, void *stream) {
    return launch_vector_to_angle(d_vec_n3, d_a, d_b, n, lonlat, degrees, d_status, S(stream));
}
int hpgb_reorder_map(const void *d_map_in, void *d_map_out, int64_t npix, int elem_bytes, int ring_to_nest, void *stream) {
    return launch_reorder_map(d_map_in, d_map_out, npix, elem_bytes, ring_to_nest, S(stream));
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
    PIPE(arrs, nside_v ? 3 : 2, launch_n2r((const int64_t *)d[0], (int64_t *)d[1], cn, nside, NSV(2), base, st, s));
}
int hpgb_host_ring_to_nest(const int64_t *ring, int64_t *nest, int64_t n, int64_t nside, const int64_t *nside_v,
                           int device, int64_t *first_bad) {
    hpgb_pipe_arr arrs[3] = {IN(ring, 8), OUT(nest, 8), IN(nside_v, 8)};
    PIPE(arrs, nside_v ? 3 : 2, launch_r2n((const int64_t *)d[0], (int64_t *)d[1], cn, nside, NSV(2), base, st, s));
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
int hpgb_host_thetaphi_to_lonlat(const double *theta, const double *phi, double *lon, double *lat, int64_t n, int degrees,
                                 int device, int64_t *first_bad) {
    hpgb_pipe_arr arrs[4] = {IN(theta, 8), IN(phi, 8), OUT(lon, 8), OUT(lat, 8)};
    PIPE(arrs, 4, launch_thetaphi_to_lonlat((const double *)d[0], (const double *)d[1], (double *)d[2], (double *)d[3], cn, degrees, st, s));
}
int hpgb_host_angle_to_vector(const double *a, const double *b, double *vec_n3, int64_t n, int lonlat, int degrees, int device,
                              int64_t *first_bad) {
    hpgb_pipe_arr arrs[3] = {IN(a, 8), IN(b, 8), OUT(vec_n3, 24)};
    PIPE(arrs, 3, launch_angle_to_vector((const double *)d[0], (const double *)d[1], (double *)d[2], cn, lonlat, degrees, base, st, s));
}
int hpgb_host_vector_to_angle(const double *vec_n3, double *a, double *b, int64_t n, int lonlat, int degrees, int device,
                              int64_t *first_bad) {
    hpgb_pipe_arr arrs[3] = {IN(vec_n3, 24), OUT(a, 8), OUT(b, 8)};
    PIPE(arrs, 3, launch_vector_to_angle((const double *)d[0], (double *)d[1], (double *)d[2], cn, lonlat, degrees, st, s));
}
// the whole map goes to ONE device (a permutation does not shard), is permuted there and comes back
int hpgb_host_reorder_map(const void *map_in, void *map_out, int64_t npix, int elem_bytes, int ring_to_nest, int device) {
    if (npix <= 0 || !map_in || !map_out || elem_bytes <= 0) return HPGB_E_BAD_ARGUMENT;
    device = hpgb_host_first_device(device);
    int prev = 0;
    cudaGetDevice(&prev);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return HPGB_E_CUDA + (int)e;
    const size_t bytes = (size_t)npix * (size_t)elem_bytes;
    char *d = nullptr;
    int rc = HPGB_OK;
    e = hpgb_dev_alloc((void **)&d, 2 * bytes + 256);
    if (e != cudaSuccess) {
        cudaSetDevice(prev);
        return HPGB_E_CUDA + (int)e;
    }
    char *d_out = d + ((bytes + 255) & ~(size_t)255);
    e = cudaMemcpyAsync(d, map_in, bytes, cudaMemcpyHostToDevice, (cudaStream_t)0);
    if (e == cudaSuccess) rc = launch_reorder_map(d, d_out, npix, elem_bytes, ring_to_nest, (cudaStream_t)0);
    if (e == cudaSuccess && rc == HPGB_OK) e = cudaMemcpyAsync(map_out, d_out, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)0);
    cudaError_t e2 = cudaStreamSynchronize((cudaStream_t)0);
    hpgb_dev_free(d);
    cudaSetDevice(prev);
    if (rc != HPGB_OK) return rc;
    if (e != cudaSuccess) return HPGB_E_CUDA + (int)e;
    return e2 == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e2;
}
int hpgb_host_boundaries(const int64_t *pix, double *a_n4s, double *b_n4s, int64_t n, int64_t nside, const int64_t *nside_v,
                         int64_t step, int nest, int lonlat, int degrees, int device, int64_t *first_bad) {
    if (step < 1) return HPGB_E_BAD_ARGUMENT;
    hpgb_pipe_arr arrs[4] = {IN(pix, 8), OUT(a_n4s, 32 * step), OUT(b_n4s, 32 * step), IN(nside_v, 8)};
    PIPE(arrs, nside_v ? 4 : 3,
         launch_boundaries((const int64_t *)d[0], (double *)d[1], (double *)d[2], cn, nside, NSV(3), step, nest, lonlat, degrees, base, st, s));
}
int hpgb_host_max_pixel_radius(const int64_t *nside, double *radius, int64_t n, int degrees, int device, int64_t *first_bad) {
    hpgb_pipe_arr arrs[2] = {IN(nside, 8), OUT(radius, 8)};
    PIPE(arrs, 2, launch_max_pixrad((const int64_t *)d[0], (double *)d[1], cn, degrees, base, st, s));
}

}  // extern "C"
