// hpgb_internal.h -- kernel launchers shared between the k_*.cu translation units and hpgb_abi.cu.
// All take DEVICE pointers, the global index `base` of element 0 (chunked host pipeline) and
// launch asynchronously on `s`; return HPGB_OK / HPGB_E_*.
#ifndef HPGB_INTERNAL_H
#define HPGB_INTERNAL_H
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hpgb.h"

int launch_n2r(const int64_t *in, int64_t *out, int64_t n, int64_t nside, const int64_t *nside_v, int64_t base,
               hpgb_status_t *st, cudaStream_t s);
int launch_r2n(const int64_t *in, int64_t *out, int64_t n, int64_t nside, const int64_t *nside_v, int64_t base,
               hpgb_status_t *st, cudaStream_t s);
int launch_a2p(const double *a, const double *b, int64_t *out, int64_t n, int64_t nside, const int64_t *nside_v,
               int nest, int lonlat, int degrees, int64_t base, hpgb_status_t *st, cudaStream_t s);
int launch_p2a(const int64_t *pix, double *a, double *b, int64_t n, int64_t nside, const int64_t *nside_v, int nest,
               int lonlat, int degrees, int64_t base, hpgb_status_t *st, cudaStream_t s);
int launch_v2p(const double *x, const double *y, const double *z, int64_t *out, int64_t n, int64_t nside,
               const int64_t *nside_v, int nest, int64_t base, hpgb_status_t *st, cudaStream_t s);
int launch_p2v(const int64_t *pix, double *x, double *y, double *z, int64_t n, int64_t nside, const int64_t *nside_v,
               int nest, int64_t base, hpgb_status_t *st, cudaStream_t s);
int launch_neighbors(const int64_t *pix, int64_t *out, int64_t n, int64_t nside, const int64_t *nside_v, int nest,
                     int64_t base, hpgb_status_t *st, cudaStream_t s);
int launch_interp(const double *a, const double *b, int64_t *pix, double *wgt, int64_t n, int64_t nside,
                  const int64_t *nside_v, int nest, int lonlat, int degrees, int64_t base, hpgb_status_t *st,
                  cudaStream_t s);
int launch_upgrade(const int64_t *pix, int64_t *out, int64_t n, int64_t nside, int64_t nside_upgrade, int nest,
                   int64_t base, hpgb_status_t *st, cudaStream_t s);
int launch_lonlat(const double *lon, const double *lat, double *theta, double *phi, int64_t n, int degrees,
                  int64_t base, hpgb_status_t *st, cudaStream_t s);
int launch_thetaphi_to_lonlat(const double *theta, const double *phi, double *lon, double *lat, int64_t n, int degrees,
                              hpgb_status_t *st, cudaStream_t s);
int launch_angle_to_vector(const double *a, const double *b, double *vec3, int64_t n, int lonlat, int degrees, int64_t base,
                           hpgb_status_t *st, cudaStream_t s);
int launch_vector_to_angle(const double *vec3, double *a, double *b, int64_t n, int lonlat, int degrees, hpgb_status_t *st,
                           cudaStream_t s);
int launch_reorder_map(const void *in, void *out, int64_t npix, int elem_bytes, int ring_to_nest, cudaStream_t s);
int launch_boundaries(const int64_t *pix, double *a, double *b, int64_t n, int64_t nside, const int64_t *nside_v,
                      int64_t step, int nest, int lonlat, int degrees, int64_t base, hpgb_status_t *st, cudaStream_t s);
int launch_max_pixrad(const int64_t *nside, double *out, int64_t n, int degrees, int64_t base, hpgb_status_t *st,
                      cudaStream_t s);
int launch_ranges_offsets(const int64_t *ranges, int64_t m, int inclusive, int64_t *offs, hpgb_status_t *st, cudaStream_t s);
int launch_ranges_expand(const int64_t *ranges, const int64_t *offs, int64_t m, int64_t out_begin, int64_t out_count,
                         int64_t *out, cudaStream_t s);
int launch_qc_ring_ranges(const hpgb_disc_t *plan, int64_t *ranges, int64_t *work, cudaStream_t s);
// stream-ordered device scratch (cudaMallocAsync / cudaFreeAsync on `s`; hpgb_host.cu)
cudaError_t hpgb_dev_alloc_on(void **p, size_t bytes, cudaStream_t s);
void hpgb_dev_free_on(void *p, cudaStream_t s);
#endif
