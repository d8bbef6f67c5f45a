/*
 * hpgb.h -- C ABI of libhpgb.so, the B200 (sm_100a) implementation of hpgeom's elementwise
 * HEALPix conversion path.
 *
 * This is the drop-in boundary.  The reference has no C ABI of its own for this path: it
 * exports 15 CPython methods (reference hpgeom/hpgeom.c:3898-3929) whose per-element loops
 * call scalar cores declared in hpgeom/healpix_geom.h:64-129.  Each entry point below replaces
 * one of those method bodies (loop + scalar core); the comment above it cites the method.
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / numpy / CUDA types in the signatures
 *    (`stream` is a cudaStream_t passed as void*; NULL = the legacy default stream);
 *  - arrays are structure-of-arrays, C-contiguous, int64 / float64, n elements; row outputs
 *    ((n,8) neighbours, (n,4) interpolation) are row-major;
 *  - `nside_v` is NULL for a scalar nside (the fast, templated path) or a device array of n
 *    per-element nside values (the reference broadcasts nside against the data);
 *  - hpgb_<fn>()      : DEVICE pointers, asynchronous on `stream`, no allocation, no sync;
 *  - hpgb_host_<fn>() : HOST pointers (pinned or pageable); runs a chunked, multi-stream
 *    H2D -> kernel -> D2H pipeline on `device` (or on every GPU of the box, HPGB_DEVICE_ALL) and returns
 *    when the outputs are complete; slots are sized by a byte budget (HPGB_HOST_SLOT_BYTES, 96 MiB);
 *  - return value: HPGB_OK, an HPGB_E_* argument error, or 1000 + cudaError_t;
 *  - element errors (the reference's ValueError checks) never trap: the kernel writes -1 / NaN
 *    for the offending element and records the SMALLEST failing element index with atomicMin
 *    in `status->first_bad` (device memory, see hpgb_status_reset), so "first bad element wins"
 *    exactly like the reference's sequential loop (hpgeom/hpgeom.c:309-316).  The caller then
 *    formats the reference's message for that one element with hpgb_explain_*().
 */
#ifndef HPGB_H
#define HPGB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPGB_OK 0
#define HPGB_E_NSIDE_NOT_POSITIVE 1 /* hpgeom/hpgeom_utils.c:30-33 */
#define HPGB_E_NSIDE_NOT_POW2 2     /* hpgeom/hpgeom_utils.c:35-38 */
#define HPGB_E_NSIDE_TOO_LARGE 3    /* hpgeom/hpgeom_utils.c:40-44 */
#define HPGB_E_BAD_ARGUMENT 4
#define HPGB_E_NO_DEVICE 5
#define HPGB_E_BAD_ANGLE 6  /* query plans: text from hpgb_explain_angle */
#define HPGB_E_BAD_RADIUS 7 /* hpgeom/hpgeom_utils.c:89-98 */
#define HPGB_E_BAD_FACT 8   /* hpgeom/hpgeom_utils.c:73-87 */
#define HPGB_E_CUDA 1000 /* + cudaError_t */

#define HPGB_NO_BAD_ELEMENT 0xFFFFFFFFFFFFFFFFull

/* `device` argument of the hpgb_host_<fn>() entry points: a CUDA device ordinal, or HPGB_DEVICE_ALL = cut the
 * arrays into contiguous shards over the device set (hpgb_host_set_devices; default every visible GPU), one host
 * thread and one pipeline per GPU, from this one process.  The split is the reference's split of its element loop
 * over worker threads (hpgeom/hpgeom.c:272-279).  No collective: the shards are independent. */
#define HPGB_DEVICE_ALL (-1)

/* element error classes reported by hpgb_explain_* (reference message templates:
 * hpgeom/hpgeom_utils.c:31,36,41,52,56,66,132,134) */
#define HPGB_BAD_NONE 0
#define HPGB_BAD_NSIDE_NOT_POSITIVE 1
#define HPGB_BAD_NSIDE_NOT_POW2 2
#define HPGB_BAD_NSIDE_TOO_LARGE 3
#define HPGB_BAD_THETA 4
#define HPGB_BAD_PHI 5
#define HPGB_BAD_LAT 6
#define HPGB_BAD_PIXEL 7

typedef struct hpgb_status {
    unsigned long long first_bad; /* smallest failing element index, or HPGB_NO_BAD_ELEMENT */
} hpgb_status_t;

/* library / device management */
int hpgb_version(void);
int hpgb_device_count(void);
const char *hpgb_error_string(int code);
/* async memset of *d_status to "no bad element" */
int hpgb_status_reset(hpgb_status_t *d_status, void *stream);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
int64_t hpgb_launch_count(void);
/* free the cached device scratch / streams used by the hpgb_host_* entry points */
int hpgb_host_release(void);
/* the GPUs HPGB_DEVICE_ALL shards over (n_dev = 0: every visible device); get returns the set's size */
int hpgb_host_set_devices(const int *devices, int n_dev);
int hpgb_host_get_devices(int *devices, int cap);
/* pinned host memory for callers that want zero staging in hpgb_host_* */
int hpgb_pinned_alloc(void **ptr, int64_t bytes);
int hpgb_pinned_free(void *ptr);
/* the parallel staging copy the hpgb_host_* entry points use for pageable buffers (host only, no GPU needed) */
int hpgb_host_copy(void *dst, const void *src, int64_t bytes);

/* ---- nest_to_ring / ring_to_nest: hpgeom/hpgeom.c:1406-1653, :1676-1927 ------------------ */
int hpgb_nest_to_ring(const int64_t *d_nest, int64_t *d_ring, int64_t n, int64_t nside,
                      const int64_t *d_nside_v, hpgb_status_t *d_status, void *stream);
int hpgb_ring_to_nest(const int64_t *d_ring, int64_t *d_nest, int64_t n, int64_t nside,
                      const int64_t *d_nside_v, hpgb_status_t *d_status, void *stream);
int hpgb_host_nest_to_ring(const int64_t *nest, int64_t *ring, int64_t n, int64_t nside,
                           const int64_t *nside_v, int device, int64_t *first_bad);
int hpgb_host_ring_to_nest(const int64_t *ring, int64_t *nest, int64_t n, int64_t nside,
                           const int64_t *nside_v, int device, int64_t *first_bad);

/* ---- angle_to_pixel: hpgeom/hpgeom.c:101-380 ---------------------------------------------- */
int hpgb_angle_to_pixel(const double *d_a, const double *d_b, int64_t *d_pix, int64_t n,
                        int64_t nside, const int64_t *d_nside_v, int nest, int lonlat, int degrees,
                        hpgb_status_t *d_status, void *stream);
int hpgb_host_angle_to_pixel(const double *a, const double *b, int64_t *pix, int64_t n,
                             int64_t nside, const int64_t *nside_v, int nest, int lonlat,
                             int degrees, int device, int64_t *first_bad);

/* ---- pixel_to_angle: hpgeom/hpgeom.c:401-676 ---------------------------------------------- */
int hpgb_pixel_to_angle(const int64_t *d_pix, double *d_a, double *d_b, int64_t n, int64_t nside,
                        const int64_t *d_nside_v, int nest, int lonlat, int degrees,
                        hpgb_status_t *d_status, void *stream);
int hpgb_host_pixel_to_angle(const int64_t *pix, double *a, double *b, int64_t n, int64_t nside,
                             const int64_t *nside_v, int nest, int lonlat, int degrees, int device,
                             int64_t *first_bad);

/* ---- vector_to_pixel / pixel_to_vector: hpgeom/hpgeom.c:2298-2566, :2592-2866 -------------- */
int hpgb_vector_to_pixel(const double *d_x, const double *d_y, const double *d_z, int64_t *d_pix,
                         int64_t n, int64_t nside, const int64_t *d_nside_v, int nest,
                         hpgb_status_t *d_status, void *stream);
int hpgb_pixel_to_vector(const int64_t *d_pix, double *d_x, double *d_y, double *d_z, int64_t n,
                         int64_t nside, const int64_t *d_nside_v, int nest,
                         hpgb_status_t *d_status, void *stream);
int hpgb_host_vector_to_pixel(const double *x, const double *y, const double *z, int64_t *pix,
                              int64_t n, int64_t nside, const int64_t *nside_v, int nest,
                              int device, int64_t *first_bad);
int hpgb_host_pixel_to_vector(const int64_t *pix, double *x, double *y, double *z, int64_t n,
                              int64_t nside, const int64_t *nside_v, int nest, int device,
                              int64_t *first_bad);

/* ---- neighbors: hpgeom/hpgeom.c:2889-3184; out is (n,8) row-major, order SW,W,NW,N,NE,E,SE,S */
int hpgb_neighbors(const int64_t *d_pix, int64_t *d_out_n8, int64_t n, int64_t nside,
                   const int64_t *d_nside_v, int nest, hpgb_status_t *d_status, void *stream);
int hpgb_host_neighbors(const int64_t *pix, int64_t *out_n8, int64_t n, int64_t nside,
                        const int64_t *nside_v, int nest, int device, int64_t *first_bad);

/* ---- get_interpolation_weights: hpgeom/hpgeom.c:3464-3777; (n,4) pixels + (n,4) weights ---- */
int hpgb_interp_weights(const double *d_a, const double *d_b, int64_t *d_pix_n4, double *d_wgt_n4,
                        int64_t n, int64_t nside, const int64_t *d_nside_v, int nest, int lonlat,
                        int degrees, hpgb_status_t *d_status, void *stream);
int hpgb_host_interp_weights(const double *a, const double *b, int64_t *pix_n4, double *wgt_n4,
                             int64_t n, int64_t nside, const int64_t *nside_v, int nest,
                             int lonlat, int degrees, int device, int64_t *first_bad);

/* ---- upgrade_pixels: hpgeom/hpgeom.py:515-556 (+ pixel_ranges_to_pixels, hpgeom.c:3795-3896,
 * for the equal-length ranges upgrade_pixels produces); out holds n*(nside_upgrade/nside)^2 ---- */
int hpgb_upgrade_pixels(const int64_t *d_pix, int64_t *d_out, int64_t n, int64_t nside,
                        int64_t nside_upgrade, int nest, hpgb_status_t *d_status, void *stream);
int hpgb_host_upgrade_pixels(const int64_t *pix, int64_t *out, int64_t n, int64_t nside,
                             int64_t nside_upgrade, int nest, int device, int64_t *first_bad);

/* ---- lonlat_to_thetaphi: hpgeom/hpgeom.py:60-88 (numpy formula, not the C one) ------------- */
int hpgb_lonlat_to_thetaphi(const double *d_lon, const double *d_lat, double *d_theta,
                            double *d_phi, int64_t n, int degrees, hpgb_status_t *d_status,
                            void *stream);
int hpgb_host_lonlat_to_thetaphi(const double *lon, const double *lat, double *theta, double *phi,
                                 int64_t n, int degrees, int device, int64_t *first_bad);

/* ---- thetaphi_to_lonlat: hpgeom/hpgeom.py:91-112 (lat = -(theta - pi/2), numpy rad2deg = x * (180/pi)); no
 * validation; bit-exact ------------------------------------------------------------------------------------- */
int hpgb_thetaphi_to_lonlat(const double *d_theta, const double *d_phi, double *d_lon, double *d_lat,
                            int64_t n, int degrees, hpgb_status_t *d_status, void *stream);
int hpgb_host_thetaphi_to_lonlat(const double *theta, const double *phi, double *lon, double *lat,
                                 int64_t n, int degrees, int device, int64_t *first_bad);

/* ---- angle_to_vector: hpgeom/hpgeom.py:364-393; vec is (n,3) row-major.  Element errors: first_bad = index of
 * the first bad latitude / colatitude; a bad longitude (theta/phi form only, checked after ALL colatitudes like the
 * reference's np.any) is reported as (1 << 62) | index --------------------------------------------------------- */
int hpgb_angle_to_vector(const double *d_a, const double *d_b, double *d_vec_n3, int64_t n, int lonlat,
                         int degrees, hpgb_status_t *d_status, void *stream);
int hpgb_host_angle_to_vector(const double *a, const double *b, double *vec_n3, int64_t n, int lonlat,
                              int degrees, int device, int64_t *first_bad);

/* ---- vector_to_angle: hpgeom/hpgeom.py:396-422; vec is (n,3) row-major, need not be normalised; no validation */
int hpgb_vector_to_angle(const double *d_vec_n3, double *d_a, double *d_b, int64_t n, int lonlat, int degrees,
                         hpgb_status_t *d_status, void *stream);
int hpgb_host_vector_to_angle(const double *vec_n3, double *a, double *b, int64_t n, int lonlat, int degrees,
                              int device, int64_t *first_bad);

/* ---- reorder: hpgeom/hpgeom.py:425-462, the full-map RING <-> NEST permutation as ONE kernel
 * (map_out[ring_to_nest(i)] = map_in[i], or the converse).  npix = 12 nside^2 elements of elem_bytes = 1, 2, 4, 8
 * or 16 bytes each, nside a power of two; map_in != map_out.  Returns HPGB_E_BAD_ARGUMENT for an illegal npix. */
int hpgb_reorder_map(const void *d_map_in, void *d_map_out, int64_t npix, int elem_bytes, int ring_to_nest,
                     void *stream);
int hpgb_host_reorder_map(const void *map_in, void *map_out, int64_t npix, int elem_bytes, int ring_to_nest,
                          int device);

/* ---- boundaries: hpgeom/hpgeom.c:1956-2278 (core hpgeom/healpix_geom.c:807-888); outputs are
 * (n, 4*step) row-major: `step` points along each of the four edges, starting at the north corner
 * and going N->W, W->S, S->E, E->N ---- */
int hpgb_boundaries(const int64_t *d_pix, double *d_a_n4s, double *d_b_n4s, int64_t n, int64_t nside,
                    const int64_t *d_nside_v, int64_t step, int nest, int lonlat, int degrees,
                    hpgb_status_t *d_status, void *stream);
int hpgb_host_boundaries(const int64_t *pix, double *a_n4s, double *b_n4s, int64_t n, int64_t nside,
                         const int64_t *nside_v, int64_t step, int nest, int lonlat, int degrees,
                         int device, int64_t *first_bad);

/* ---- max_pixel_radius: hpgeom/hpgeom.c:3202-3440 (core hpgeom/healpix_geom.c:480-502); one value
 * per nside (RING rules: any positive nside <= 2^29) ---- */
int hpgb_max_pixel_radius(const int64_t *d_nside, double *d_radius, int64_t n, int degrees,
                          hpgb_status_t *d_status, void *stream);
int hpgb_host_max_pixel_radius(const int64_t *nside, double *radius, int64_t n, int degrees, int device,
                               int64_t *first_bad);

/* ---- pixel_ranges_to_pixels: hpgeom/hpgeom.c:3795-3896.  Ranges are (m,2) row-major [lo, hi) rows
 * (16-byte aligned); `inclusive` adds 1 to every hi.  Like the reference's two loops, two calls:
 *   hpgb_ranges_offsets : the count loop (:3843-3856).  Flags the first row with hi < lo in *d_status
 *                         and writes the exclusive prefix sums of the lengths to d_offsets_mp1[0..m]
 *                         (d_offsets_mp1[m] = number of output pixels; the caller reads it to allocate);
 *   hpgb_ranges_expand  : the expand loop (:3868-3876) for the window [out_begin, out_begin+out_count)
 *                         of the pixel list, written to d_out[0..out_count).
 * Host forms: count returns the output size, expand fills `out` (total pixels) through the chunked
 * D2H pipeline. ---- */
int hpgb_ranges_offsets(const int64_t *d_ranges_m2, int64_t m, int inclusive, int64_t *d_offsets_mp1,
                        hpgb_status_t *d_status, void *stream);
int hpgb_ranges_expand(const int64_t *d_ranges_m2, const int64_t *d_offsets_mp1, int64_t m,
                       int64_t out_begin, int64_t out_count, int64_t *d_out, void *stream);
int hpgb_host_ranges_count(const int64_t *ranges_m2, int64_t m, int inclusive, int device,
                           int64_t *total, int64_t *first_bad);
int hpgb_host_ranges_expand(const int64_t *ranges_m2, int64_t m, int inclusive, int64_t *out,
                            int64_t total, int device);

/* ---- query_circle, RING scheme: hpgeom/hpgeom.c:765-864 -> query_disc, hpgeom/healpix_geom.c:632-747.
 * The reference walks the rings the disc touches one by one; here every ring is one GPU thread.
 *   hpgb_query_circle_ring_plan   : host only, O(1): the wrapper's argument checks in its order
 *                                   (angles, radius, nside, fact: hpgeom/hpgeom.c:793-838; returns
 *                                   HPGB_E_BAD_ANGLE / _RADIUS / nside codes / _FACT) and the per-disc
 *                                   scalars of :638-674 and :737-743.  inclusive = 0 ignores fact.
 *                                   plan->m = rows of the range table.
 *   hpgb_query_circle_ring_ranges : fills the (m,2) table of [lo, hi) pixel intervals (ascending,
 *                                   disjoint, unused rows (0,0)): row 0 = north polar range, rows
 *                                   1+2j / 2+2j = ring irmin+j, last row = south polar range.
 *                                   d_work = hpgb_query_circle_ring_work_bytes(plan) bytes of device
 *                                   scratch (16-byte aligned; 0 bytes / NULL unless fact > 1): the
 *                                   inclusive edge scans the reference does sequentially (:701-710)
 *                                   are spread over the whole grid there.
 * The pixel list is then hpgb_ranges_offsets + hpgb_ranges_expand on that table.
 * Host forms: count returns the number of pixels, fill writes them (host buffer of `total`). ---- */
typedef struct hpgb_disc {
    int64_t nside, fct;        /* fct = max(fact, 1) */
    int64_t whole_sphere;      /* 1: the disc covers every pixel; the table is the single row (0, npix) */
    int64_t irmin, irmax;      /* rings scanned (irmin > irmax: none) */
    int64_t north_hi;          /* > 0: the polar range [0, north_hi) */
    int64_t south_lo;          /* >= 0: the polar range [south_lo, npix) */
    int64_t cpix;              /* RING pixel holding the disc centre */
    int64_t m;                 /* rows of the range table */
    double phi, z0, xa, cosrbig, cosrsmall;
} hpgb_disc_t;
int hpgb_query_circle_ring_plan(int64_t nside, double a, double b, double radius, int inclusive,
                                int64_t fact, int lonlat, int degrees, hpgb_disc_t *plan);
int64_t hpgb_query_circle_ring_work_bytes(const hpgb_disc_t *plan);
int hpgb_query_circle_ring_ranges(const hpgb_disc_t *plan, int64_t *d_ranges_m2, void *d_work, void *stream);
int hpgb_host_query_circle_ring_count(const hpgb_disc_t *plan, int device, int64_t *total);
int hpgb_host_query_circle_ring_fill(const hpgb_disc_t *plan, int64_t *out, int64_t total, int device);

/* ---- many discs at once, RING scheme (an extension: the reference has no batched call).  One small disc is
 * launch-latency bound on a GPU, a batch is not: D plans of hpgb_query_circle_ring_plan (same nside and fact)
 * are prepared on the host, ONE launch runs every (disc, ring) pair and one offsets + expand pass writes the
 * D pixel lists back to back.  prepare: normalises the plans and fills ring_start[0..D] (returns the total ring
 * count); batch_ranges: device form, table of 2 * rings + 2 * D rows = the D single-disc tables concatenated;
 * host form: call with out = NULL for disc_offsets[0..D] and *total, then again with the buffer. ---- */
int hpgb_query_circle_ring_plan_many(int64_t nside, const double *a, const double *b, const double *radius,
                                     int64_t n_discs, int inclusive, int64_t fact, int lonlat, int degrees,
                                     hpgb_disc_t *plans, int64_t *first_bad);
int64_t hpgb_query_circle_ring_batch_prepare(hpgb_disc_t *plans, int64_t n_discs, int64_t *ring_start);
int hpgb_query_circle_ring_batch_ranges(const hpgb_disc_t *d_plans, const int64_t *d_ring_start, int64_t n_discs,
                                        int64_t total_rings, int64_t nside, int64_t fct, int64_t *d_ranges_m2,
                                        void *d_work, void *stream);
int hpgb_host_query_circle_ring_batch(hpgb_disc_t *plans, int64_t n_discs, int device, int64_t *disc_offsets,
                                      int64_t *total, int64_t *out, int64_t cap);

/* ---- query_circle, NEST scheme: the hierarchical search of query_disc (hpgeom/healpix_geom.c:748-800,
 * check_pixel_nest :541-601), walked breadth first: one kernel launch per order over the whole frontier,
 * the emitted ranges sorted on the device.  Frontier sizes are data dependent, so this entry point owns its
 * device memory: open() runs the search on `device` (same argument checks and return codes as the RING
 * plan; fact must be a power of two) and keeps the sorted range table + offsets in an opaque handle;
 * m = table rows, total = pixels.  Then: the merged range list (return_pixel_ranges=True), the pixel
 * list into a host buffer (chunked D2H pipeline) or into a device buffer on `stream`; close() frees. ---- */
int hpgb_query_circle_nest_open(int64_t nside, double a, double b, double radius, int inclusive, int64_t fact,
                                int lonlat, int degrees, int device, void **handle, int64_t *m, int64_t *total);
int hpgb_query_nest_ranges(void *handle, int64_t *out_m2, int64_t *n_merged);
int hpgb_query_nest_pixels_host(void *handle, int64_t *out, int64_t total);
int hpgb_query_nest_pixels_device(void *handle, int64_t *d_out, int64_t total, void *stream);
int hpgb_query_nest_close(void *handle);

/* ---- many discs at once, NEST scheme (extension): the same search over a frontier of (disc, pixel) pairs, every
 * order ONE launch for the whole batch, ranges ordered by (disc, lo) with two stable radix sorts.  The handle is
 * used like a single-disc one (hpgb_query_nest_pixels_host / _device, hpgb_query_nest_close); the D + 1
 * boundaries of the pixel lists come from hpgb_query_nest_batch_offsets.  first_bad = the first disc whose
 * arguments the reference would reject (the return code says why). ---- */
int hpgb_query_circle_nest_batch_open(int64_t nside, const double *a, const double *b, const double *radius,
                                      int64_t n_discs, int inclusive, int64_t fact, int lonlat, int degrees,
                                      int device, void **handle, int64_t *m, int64_t *total, int64_t *first_bad);
int hpgb_query_nest_batch_offsets(void *handle, int64_t *disc_offsets);

/* ---- error text for ONE element (host side, pure functions) -------------------------------
 * Re-evaluates the reference's checks in the reference's order for the element the kernel
 * flagged and writes the reference's message (at most `cap` bytes, NUL terminated).
 * Returns the HPGB_BAD_* class (HPGB_BAD_NONE if the element is in fact valid). */
int hpgb_explain_nside(int64_t nside, int nest, char *msg, int cap);
int hpgb_explain_angle(int64_t nside, int nest, double a, double b, int lonlat, int degrees,
                       char *msg, int cap);
int hpgb_explain_pixel(int64_t nside, int nest, int64_t pix, char *msg, int cap);

#ifdef __cplusplus
}
#endif
#endif /* HPGB_H */
