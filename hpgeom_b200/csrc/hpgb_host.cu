// hpgb_host.cu -- host-side runtime of libhpgb.so: device bookkeeping, the chunked
// H2D -> kernel -> D2H pipeline behind the hpgb_host_* entry points, pinned memory helpers and
// the reference-compatible error text for ONE flagged element.
//
// Nothing in this file converts data: every element goes through the CUDA kernels of
// hpgb_kernels.cu.  The only arithmetic here is hpgb_explain_*(), which re-evaluates the
// reference's range checks for the single element a kernel flagged so that the message matches
// the reference's (hpgeom/hpgeom_utils.c:29-71,119-143) character for character.
#include <cuda_runtime.h>
#include <inttypes.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "hpgb_core.h"
#include "hpgb_host.h"

std::atomic<int64_t> g_hpgb_launches{0};

int hpgb_sm_count() {
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!sms[dev]) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        sms[dev] = v > 0 ? v : 148;
    }
    return sms[dev];
}

extern "C" int hpgb_version(void) { return 100; }

extern "C" int hpgb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" const char *hpgb_error_string(int code) {
    switch (code) {
        case HPGB_OK: return "ok";
        case HPGB_E_NSIDE_NOT_POSITIVE: return "nside must be positive";
        case HPGB_E_NSIDE_NOT_POW2: return "nside must be power of 2 for NEST pixels";
        case HPGB_E_NSIDE_TOO_LARGE: return "nside must not be greater than 2**29";
        case HPGB_E_BAD_ARGUMENT: return "bad argument";
        case HPGB_E_NO_DEVICE: return "no CUDA device";
        default: break;
    }
    if (code >= HPGB_E_CUDA) return cudaGetErrorString((cudaError_t)(code - HPGB_E_CUDA));
    return "unknown error";
}

extern "C" int hpgb_status_reset(hpgb_status_t *d_status, void *stream) {
    cudaError_t e = cudaMemsetAsync(d_status, 0xFF, sizeof(hpgb_status_t), (cudaStream_t)stream);
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

extern "C" int64_t hpgb_launch_count(void) { return g_hpgb_launches.load(); }

extern "C" int hpgb_pinned_alloc(void **ptr, int64_t bytes) {
    cudaError_t e = cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocPortable);
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

extern "C" int hpgb_pinned_free(void *ptr) {
    cudaError_t e = cudaFreeHost(ptr);
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

// ---------------------------------------------------------------------------------------------
// chunked host pipeline
// ---------------------------------------------------------------------------------------------
namespace {

struct HostCtx {
    bool ready = false;
    cudaStream_t stream[HPGB_PIPE_SLOTS];
    char *dbuf[HPGB_PIPE_SLOTS] = {nullptr, nullptr, nullptr};
    size_t dbytes[HPGB_PIPE_SLOTS] = {0, 0, 0};
    hpgb_status_t *d_status = nullptr;
    hpgb_status_t *h_status = nullptr;
};

HostCtx g_ctx[64];
std::mutex g_ctx_mutex;

int64_t chunk_points() {
    static int64_t v = 0;
    if (!v) {
        const char *e = getenv("HPGB_HOST_CHUNK");
        v = e ? atoll(e) : ((int64_t)1 << 22);
        if (v < 2) v = 2;
        v &= ~(int64_t)1;
    }
    return v;
}

#define CK(call)                                            \
    do {                                                    \
        cudaError_t e_ = (call);                            \
        if (e_ != cudaSuccess) return HPGB_E_CUDA + (int)e_; \
    } while (0)

}  // namespace

int hpgb_host_pipeline(int device, int64_t n, const hpgb_pipe_arr *arrs, int narr,
                       const hpgb_pipe_launch &launch, int64_t *first_bad) {
    if (first_bad) *first_bad = -1;
    if (n <= 0) return HPGB_OK;
    int ndev = hpgb_device_count();
    if (ndev <= 0) return HPGB_E_NO_DEVICE;
    if (device < 0 || device >= ndev || device >= 64 || narr > HPGB_PIPE_MAX_ARR) return HPGB_E_BAD_ARGUMENT;

    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    int prev = 0;
    cudaGetDevice(&prev);
    CK(cudaSetDevice(device));
    HostCtx &c = g_ctx[device];
    if (!c.ready) {
        for (int s = 0; s < HPGB_PIPE_SLOTS; s++) CK(cudaStreamCreateWithFlags(&c.stream[s], cudaStreamNonBlocking));
        CK(cudaMalloc(&c.d_status, sizeof(hpgb_status_t)));
        CK(cudaHostAlloc(&c.h_status, sizeof(hpgb_status_t), cudaHostAllocDefault));
        c.ready = true;
    }

    const int64_t cp = chunk_points() < n ? chunk_points() : ((n + 1) & ~(int64_t)1);
    size_t off[HPGB_PIPE_MAX_ARR + 1];
    size_t total = 0;
    for (int a = 0; a < narr; a++) {
        off[a] = total;
        total += (((size_t)cp * (size_t)arrs[a].bytes_per_point) + 255) & ~(size_t)255;
    }
    for (int s = 0; s < HPGB_PIPE_SLOTS; s++) {
        if (c.dbytes[s] < total) {
            if (c.dbuf[s]) CK(cudaFree(c.dbuf[s]));
            c.dbuf[s] = nullptr;
            c.dbytes[s] = 0;
            CK(cudaMalloc(&c.dbuf[s], total));
            c.dbytes[s] = total;
        }
    }
    CK(cudaMemsetAsync(c.d_status, 0xFF, sizeof(hpgb_status_t), c.stream[0]));
    CK(cudaStreamSynchronize(c.stream[0]));

    int rc = HPGB_OK;
    int slot = 0;
    for (int64_t base = 0; base < n && rc == HPGB_OK; base += cp, slot = (slot + 1) % HPGB_PIPE_SLOTS) {
        const int64_t cn = (n - base) < cp ? (n - base) : cp;
        cudaStream_t st = c.stream[slot];
        void *dptr[HPGB_PIPE_MAX_ARR];
        for (int a = 0; a < narr; a++) {
            dptr[a] = c.dbuf[slot] + off[a];
            if (arrs[a].h_in)
                CK(cudaMemcpyAsync(dptr[a], arrs[a].h_in + (size_t)base * arrs[a].bytes_per_point,
                                   (size_t)cn * arrs[a].bytes_per_point, cudaMemcpyHostToDevice, st));
        }
        rc = launch(dptr, base, cn, c.d_status, st);
        if (rc != HPGB_OK) break;
        for (int a = 0; a < narr; a++) {
            if (arrs[a].h_out)
                CK(cudaMemcpyAsync(arrs[a].h_out + (size_t)base * arrs[a].bytes_per_point, dptr[a],
                                   (size_t)cn * arrs[a].bytes_per_point, cudaMemcpyDeviceToHost, st));
        }
    }
    for (int s = 0; s < HPGB_PIPE_SLOTS; s++) {
        cudaError_t e = cudaStreamSynchronize(c.stream[s]);
        if (e != cudaSuccess && rc == HPGB_OK) rc = HPGB_E_CUDA + (int)e;
    }
    if (rc == HPGB_OK) {
        CK(cudaMemcpy(c.h_status, c.d_status, sizeof(hpgb_status_t), cudaMemcpyDeviceToHost));
        if (first_bad && c.h_status->first_bad != HPGB_NO_BAD_ELEMENT) *first_bad = (int64_t)c.h_status->first_bad;
    }
    cudaSetDevice(prev);
    return rc;
}

extern "C" int hpgb_host_release(void) {
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    int prev = 0;
    cudaGetDevice(&prev);
    for (int d = 0; d < 64; d++) {
        HostCtx &c = g_ctx[d];
        if (!c.ready) continue;
        cudaSetDevice(d);
        for (int s = 0; s < HPGB_PIPE_SLOTS; s++) {
            cudaStreamDestroy(c.stream[s]);
            if (c.dbuf[s]) cudaFree(c.dbuf[s]);
            c.dbuf[s] = nullptr;
            c.dbytes[s] = 0;
        }
        cudaFree(c.d_status);
        cudaFreeHost(c.h_status);
        c.ready = false;
    }
    cudaSetDevice(prev);
    return HPGB_OK;
}

// ---------------------------------------------------------------------------------------------
// error text (reference message templates: hpgeom/hpgeom_utils.c:31,36,41,52,56,66,132,134)
// ---------------------------------------------------------------------------------------------
extern "C" int hpgb_explain_nside(int64_t nside, int nest, char *msg, int cap) {
    int c = hpgb_check_nside(nside, nest);
    if (msg && cap > 0) {
        msg[0] = 0;
        if (c == 1) snprintf(msg, cap, "nside %" PRId64 " must be positive.", nside);
        if (c == 2) snprintf(msg, cap, "nside %" PRId64 " must be power of 2 for NEST pixels", nside);
        if (c == 3) snprintf(msg, cap, "nside %" PRId64 " must not be greater than 2**%d", nside, 29);
    }
    return c;
}

extern "C" int hpgb_explain_angle(int64_t nside, int nest, double a, double b, int lonlat, int degrees,
                                  char *msg, int cap) {
    int c = hpgb_explain_nside(nside, nest, msg, cap);
    if (c) return c;
    const double PI = 3.141592653589793238462643383279502884197;
    const double HALFPI = 1.570796326794896619231321691639751442099;
    const double TWOPI = 6.283185307179586476925286766559005768394;
    if (lonlat) {
        volatile double lat = b;
        if (degrees) {
            lat = lat * PI;  // (lat*pi)/180, two roundings like the reference's macro expansion
            lat = lat / 180.0;
        }
        if (lat < -HALFPI || lat > HALFPI) {
            if (msg && cap > 0) {
                if (degrees) {
                    volatile double back = lat * 180.0;  // lat * 180.0 / pi
                    back = back / PI;
                    snprintf(msg, cap, "lat = %g out of range [-90, 90]", (double)back);
                } else {
                    snprintf(msg, cap, "lat = %g out of range [-pi/2, pi/2]", (double)lat);
                }
            }
            return HPGB_BAD_LAT;
        }
        return HPGB_BAD_NONE;
    }
    if (a < 0.0 || a > PI) {
        if (msg && cap > 0) snprintf(msg, cap, "colatitude (theta) = %g out of range [0, pi]", a);
        return HPGB_BAD_THETA;
    }
    if (b < 0 || b > TWOPI) {
        if (msg && cap > 0) snprintf(msg, cap, "longitude (phi) = %g out of range [0, 2*pi]", b);
        return HPGB_BAD_PHI;
    }
    return HPGB_BAD_NONE;
}

extern "C" int hpgb_explain_pixel(int64_t nside, int nest, int64_t pix, char *msg, int cap) {
    int c = hpgb_explain_nside(nside, nest, msg, cap);
    if (c) return c;
    if (pix < 0 || pix >= 12 * nside * nside) {
        if (msg && cap > 0)
            snprintf(msg, cap, "Pixel value %" PRId64 " out of range for nside %" PRId64, pix, nside);
        return HPGB_BAD_PIXEL;
    }
    return HPGB_BAD_NONE;
}
