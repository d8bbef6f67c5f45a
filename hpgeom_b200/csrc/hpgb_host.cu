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
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

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

cudaError_t hpgb_dev_alloc(void **p, size_t bytes) {
    static bool tuned[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !tuned[dev]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        tuned[dev] = true;
    }
    return cudaMallocAsync(p, bytes ? bytes : 16, (cudaStream_t)0);
}
void hpgb_dev_free(void *p) {
    if (!p) return;
    int prev = 0, dev = 0;
    cudaGetDevice(&prev);
    dev = prev;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice) dev = a.device;
    if (dev != prev) cudaSetDevice(dev);  // the legacy stream is per device
    cudaFreeAsync(p, (cudaStream_t)0);
    if (dev != prev) cudaSetDevice(prev);
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
    char *hbuf[HPGB_PIPE_SLOTS] = {nullptr, nullptr, nullptr};  // pinned staging for pageable caller buffers
    size_t hbytes[HPGB_PIPE_SLOTS] = {0, 0, 0};
    cudaEvent_t done[HPGB_PIPE_SLOTS];
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

// Plain numpy arrays are pageable: cudaMemcpyAsync then goes through the driver's single-threaded
// staging (~8 GB/s measured: the drop-in ran no faster than the 16-thread reference).  For such operands
// the pipeline stages through its own pinned buffers, filled / drained by a small pool of host threads
// (parallel memcpy), so the PCIe engines see pinned memory and the host copies of chunk i+1 overlap the
// DMA of chunk i.
class CopyPool {
  public:
    static CopyPool &get() {
        static CopyPool *p = new CopyPool;  // never destroyed: its threads sleep on the condition variable at exit
        return *p;
    }
    // dst[0..bytes) = src[0..bytes), split over the pool (the caller takes the first share).  One job at a
    // time: the pipeline calls it under g_ctx_mutex.
    void copy(char *dst, const char *src, size_t bytes) {
        const size_t min_part = (size_t)1 << 20;
        int parts = (int)((bytes + min_part - 1) / min_part);
        if (parts > nworkers_ + 1) parts = nworkers_ + 1;
        if (parts <= 1) {
            memcpy(dst, src, bytes);
            return;
        }
        const size_t per = ((bytes + parts - 1) / parts + 63) & ~(size_t)63;
        {
            std::lock_guard<std::mutex> lk(m_);
            dst_ = dst, src_ = src, bytes_ = bytes, per_ = per;
            next_ = 1, parts_ = parts, pending_ = parts - 1;
        }
        cv_.notify_all();
        memcpy(dst, src, per < bytes ? per : bytes);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return pending_ == 0; });
        parts_ = 0, next_ = 0;
    }

  private:
    CopyPool() {
        const char *e = getenv("HPGB_HOST_THREADS");
        int n = e ? atoi(e) : (int)std::thread::hardware_concurrency();  // HPGB_HOST_THREADS overrides
        if (!e && n > 16) n = 16;
        if (n < 1) n = 1;
        nworkers_ = n - 1;  // the caller is one of the copiers
        for (int i = 0; i < nworkers_; i++) std::thread([this] { worker(); }).detach();
    }
    void worker() {
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            cv_.wait(lk, [&] { return next_ < parts_; });
            const int part = next_++;
            char *d = dst_;
            const char *s = src_;
            const size_t per = per_, bytes = bytes_;
            lk.unlock();
            const size_t off = (size_t)part * per;
            if (off < bytes) memcpy(d + off, s + off, (off + per <= bytes) ? per : bytes - off);
            lk.lock();
            if (--pending_ == 0) done_.notify_all();
        }
    }
    std::mutex m_;
    std::condition_variable cv_, done_;
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    size_t bytes_ = 0, per_ = 0;
    int next_ = 0, parts_ = 0, pending_ = 0, nworkers_ = 0;
};

bool is_pageable(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
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
        for (int s = 0; s < HPGB_PIPE_SLOTS; s++) {
            CK(cudaStreamCreateWithFlags(&c.stream[s], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&c.done[s], cudaEventDisableTiming));
        }
        CK(cudaMalloc(&c.d_status, sizeof(hpgb_status_t)));
        CK(cudaHostAlloc(&c.h_status, sizeof(hpgb_status_t), cudaHostAllocDefault));
        c.ready = true;
    }

    const int64_t cp = chunk_points() < n ? chunk_points() : ((n + 1) & ~(int64_t)1);
    size_t off[HPGB_PIPE_MAX_ARR + 1];
    size_t total = 0;
    bool stage[HPGB_PIPE_MAX_ARR];
    bool any_stage = false;
    for (int a = 0; a < narr; a++) {
        off[a] = total;
        total += (((size_t)cp * (size_t)arrs[a].bytes_per_point) + 255) & ~(size_t)255;
        stage[a] = is_pageable(arrs[a].h_in ? (const void *)arrs[a].h_in : (const void *)arrs[a].h_out);
        any_stage |= stage[a];
    }
    for (int s = 0; s < HPGB_PIPE_SLOTS; s++) {
        if (c.dbytes[s] < total) {
            if (c.dbuf[s]) CK(cudaFree(c.dbuf[s]));
            c.dbuf[s] = nullptr;
            c.dbytes[s] = 0;
            CK(cudaMalloc(&c.dbuf[s], total));
            c.dbytes[s] = total;
        }
        if (any_stage && c.hbytes[s] < total) {
            if (c.hbuf[s]) CK(cudaFreeHost(c.hbuf[s]));
            c.hbuf[s] = nullptr;
            c.hbytes[s] = 0;
            CK(cudaHostAlloc(&c.hbuf[s], total, cudaHostAllocDefault));
            c.hbytes[s] = total;
        }
    }
    CK(cudaMemsetAsync(c.d_status, 0xFF, sizeof(hpgb_status_t), c.stream[0]));
    CK(cudaStreamSynchronize(c.stream[0]));

    // chunk i's copy-out, chunk i+1's kernel and chunk i+2's copy-in overlap (one stream per slot).  Pageable
    // operands pass through the slot's pinned staging buffer: filled by the copy pool before the H2D is
    // queued, drained (after the slot's event) when the slot comes round again or at the end.
    int rc = HPGB_OK;
    int slot = 0;
    int64_t slot_base[HPGB_PIPE_SLOTS], slot_n[HPGB_PIPE_SLOTS];
    for (int s = 0; s < HPGB_PIPE_SLOTS; s++) slot_n[s] = 0;
    auto drain = [&](int s) -> int {  // copy the staged outputs of the chunk in slot s to the caller
        if (slot_n[s] == 0) return HPGB_OK;
        cudaError_t e = cudaEventSynchronize(c.done[s]);
        if (e != cudaSuccess) return HPGB_E_CUDA + (int)e;
        for (int a = 0; a < narr; a++)
            if (arrs[a].h_out && stage[a])
                CopyPool::get().copy(arrs[a].h_out + (size_t)slot_base[s] * arrs[a].bytes_per_point, c.hbuf[s] + off[a],
                                     (size_t)slot_n[s] * arrs[a].bytes_per_point);
        slot_n[s] = 0;
        return HPGB_OK;
    };
    for (int64_t base = 0; base < n && rc == HPGB_OK; base += cp, slot = (slot + 1) % HPGB_PIPE_SLOTS) {
        const int64_t cn = (n - base) < cp ? (n - base) : cp;
        cudaStream_t st = c.stream[slot];
        if (any_stage) {
            rc = drain(slot);
            if (rc != HPGB_OK) break;
        }
        void *dptr[HPGB_PIPE_MAX_ARR];
        for (int a = 0; a < narr; a++) {
            dptr[a] = c.dbuf[slot] + off[a];
            if (arrs[a].h_in) {
                const char *src = arrs[a].h_in + (size_t)base * arrs[a].bytes_per_point;
                const size_t bytes = (size_t)cn * arrs[a].bytes_per_point;
                if (stage[a]) {
                    CopyPool::get().copy(c.hbuf[slot] + off[a], src, bytes);
                    src = c.hbuf[slot] + off[a];
                }
                CK(cudaMemcpyAsync(dptr[a], src, bytes, cudaMemcpyHostToDevice, st));
            }
        }
        rc = launch(dptr, base, cn, c.d_status, st);
        if (rc != HPGB_OK) break;
        for (int a = 0; a < narr; a++) {
            if (arrs[a].h_out) {
                char *dst = stage[a] ? c.hbuf[slot] + off[a] : arrs[a].h_out + (size_t)base * arrs[a].bytes_per_point;
                CK(cudaMemcpyAsync(dst, dptr[a], (size_t)cn * arrs[a].bytes_per_point, cudaMemcpyDeviceToHost, st));
            }
        }
        if (any_stage) {
            CK(cudaEventRecord(c.done[slot], st));
            slot_base[slot] = base;
            slot_n[slot] = cn;
        }
    }
    if (any_stage && rc == HPGB_OK) {
        for (int k = 0; k < HPGB_PIPE_SLOTS && rc == HPGB_OK; k++) rc = drain((slot + k) % HPGB_PIPE_SLOTS);
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
            cudaEventDestroy(c.done[s]);
            if (c.dbuf[s]) cudaFree(c.dbuf[s]);
            c.dbuf[s] = nullptr;
            c.dbytes[s] = 0;
            if (c.hbuf[s]) cudaFreeHost(c.hbuf[s]);
            c.hbuf[s] = nullptr;
            c.hbytes[s] = 0;
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
