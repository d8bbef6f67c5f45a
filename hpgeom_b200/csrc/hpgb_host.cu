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
#include <chrono>
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

static void hpgb_tune_pool() {
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
}
cudaError_t hpgb_dev_alloc(void **p, size_t bytes) {
    hpgb_tune_pool();
    return cudaMallocAsync(p, bytes ? bytes : 16, (cudaStream_t)0);
}
// scratch that lives between two launches on the caller's stream (per-launch ring tables)
cudaError_t hpgb_dev_alloc_on(void **p, size_t bytes, cudaStream_t s) {
    hpgb_tune_pool();
    return cudaMallocAsync(p, bytes ? bytes : 16, s);
}
void hpgb_dev_free_on(void *p, cudaStream_t s) {
    if (p) cudaFreeAsync(p, s);
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
std::mutex g_ctx_mutex[64];  // one pipeline per device at a time; different devices run concurrently (hpgb_host_*_multi)

// Chunk size of the pipeline, in points.  The slots are sized in BYTES, not points: row-valued operands
// (upgrade_pixels: 8 * 4^k bytes per point, boundaries: 64 * step) would otherwise ask for 3 device + 3 pinned
// buffers of many GB each where the reference only needs the output array.  A slot holds at most
// HPGB_HOST_SLOT_BYTES (default 96 MiB, the size the 2^22-point chunks of angle_to_pixel use: large enough to run
// PCIe at full rate, small enough that three slots overlap on a 10M-point call); HPGB_HOST_CHUNK caps the points.
int64_t chunk_points(size_t bytes_per_point) {
    static int64_t cap = 0, budget = 0;
    if (!cap) {
        const char *e = getenv("HPGB_HOST_CHUNK");
        int64_t v = e ? atoll(e) : ((int64_t)1 << 22);
        const char *b = getenv("HPGB_HOST_SLOT_BYTES");
        int64_t w = b ? atoll(b) : ((int64_t)96 << 20);
        budget = w < 4096 ? 4096 : w;
        cap = v < 2 ? 2 : v;
    }
    int64_t v = cap;
    if (bytes_per_point > 0 && (int64_t)bytes_per_point * v > budget) v = budget / (int64_t)bytes_per_point;
    if (v < 2) v = 2;
    return v & ~(int64_t)1;
}

// Plain numpy arrays are pageable: cudaMemcpyAsync then goes through the driver's single-threaded
// staging (~8 GB/s measured: the drop-in ran no faster than the 16-thread reference).  For such operands
// the pipeline stages through its own pinned buffers, filled / drained by a small pool of host threads
// (parallel memcpy), so the PCIe engines see pinned memory and the host copies of chunk i+1 overlap the
// DMA of chunk i.
// Workers and the submitting thread SPIN for a short while (HPGB_HOST_SPIN_US, default 500 us) before they sleep on a
// condition variable: a pipeline posts three copy jobs per chunk a few hundred microseconds apart, and waking fifteen
// sleeping threads (futex + a scheduler round trip, slow inside a VM) cost ~0.4 ms per job -- 1.3 ms per chunk, as much
// as the chunk's DMA (measured: a 10M-point numpy call got SLOWER when cut into more chunks).  Only callers that pass
// pageable buffers ever post a job, so processes that use pinned buffers never spin.
class CopyPool {
  public:
    static CopyPool &get() {
        static CopyPool *p = new CopyPool;  // never destroyed: its threads sleep on the condition variable at exit
        return *p;
    }
    // dst[0..bytes) = src[0..bytes), split over the pool (the caller copies shares too).  One job at a
    // time (submit_): pipelines of different devices take turns, each job uses the whole pool.
    void copy(char *dst, const char *src, size_t bytes) {
        std::lock_guard<std::mutex> job(submit_);
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
            avail_.store(true, std::memory_order_release);
        }
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
        memcpy(dst, src, per < bytes ? per : bytes);
        std::unique_lock<std::mutex> lk(m_);
        while (next_ < parts_) {  // shares nobody has taken yet: the submitting thread helps
            const int part = take_();
            lk.unlock();
            share_(part, dst, src, per, bytes);
            lk.lock();
            --pending_;
        }
        if (pending_ != 0) {  // the last shares are in flight on other threads: they take < 100 us
            lk.unlock();
            const auto t0 = std::chrono::steady_clock::now();
            while (!done_flag_.load(std::memory_order_acquire) &&
                   std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(spin_us_))
                relax_();
            lk.lock();
            done_.wait(lk, [&] { return pending_ == 0; });
        }
        parts_ = 0, next_ = 0;
        done_flag_.store(false, std::memory_order_relaxed);
    }

  private:
    CopyPool() {
        const char *e = getenv("HPGB_HOST_THREADS");
        int n = e ? atoi(e) : (int)std::thread::hardware_concurrency();  // HPGB_HOST_THREADS overrides
        if (!e && n > 16) n = 16;
        if (n < 1) n = 1;
        const char *sp = getenv("HPGB_HOST_SPIN_US");
        spin_us_ = sp ? atoi(sp) : 500;
        if (spin_us_ < 0) spin_us_ = 0;
        nworkers_ = n - 1;  // the caller is one of the copiers
        for (int i = 0; i < nworkers_; i++) std::thread([this] { worker(); }).detach();
    }
    static void relax_() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    int take_() {  // m_ held, next_ < parts_
        const int part = next_++;
        if (next_ >= parts_) avail_.store(false, std::memory_order_release);
        return part;
    }
    static void share_(int part, char *d, const char *s, size_t per, size_t bytes) {
        const size_t off = (size_t)part * per;
        if (off < bytes) memcpy(d + off, s + off, (off + per <= bytes) ? per : bytes - off);
    }
    void worker() {
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            if (!(next_ < parts_)) {
                lk.unlock();
                bool got = false;
                if (spin_us_ > 0) {
                    const auto t0 = std::chrono::steady_clock::now();
                    do {
                        for (int k = 0; k < 64 && !got; k++) {
                            got = avail_.load(std::memory_order_acquire);
                            if (!got) relax_();
                        }
                    } while (!got && std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(spin_us_));
                }
                lk.lock();
                if (!got) {
                    sleepers_.fetch_add(1, std::memory_order_acq_rel);
                    cv_.wait(lk, [&] { return next_ < parts_; });
                    sleepers_.fetch_sub(1, std::memory_order_acq_rel);
                } else if (!(next_ < parts_)) {
                    continue;  // somebody else took the share
                }
            }
            const int part = take_();
            char *d = dst_;
            const char *s = src_;
            const size_t per = per_, bytes = bytes_;
            lk.unlock();
            share_(part, d, s, per, bytes);
            lk.lock();
            if (--pending_ == 0) {
                done_flag_.store(true, std::memory_order_release);
                done_.notify_all();
            }
        }
    }
    std::mutex m_, submit_;
    std::condition_variable cv_, done_;
    std::atomic<bool> avail_{false}, done_flag_{false};
    std::atomic<int> sleepers_{0};
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    size_t bytes_ = 0, per_ = 0;
    int next_ = 0, parts_ = 0, pending_ = 0, nworkers_ = 0, spin_us_ = 500;
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

static int pipeline_one(int device, int64_t n, const hpgb_pipe_arr *arrs, int narr, const hpgb_pipe_launch &launch,
                        int64_t *first_bad) {
    if (first_bad) *first_bad = -1;
    if (n <= 0) return HPGB_OK;
    int ndev = hpgb_device_count();
    if (ndev <= 0) return HPGB_E_NO_DEVICE;
    if (device < 0 || device >= ndev || device >= 64 || narr > HPGB_PIPE_MAX_ARR) return HPGB_E_BAD_ARGUMENT;

    std::lock_guard<std::mutex> lock(g_ctx_mutex[device]);
    struct DeviceGuard {  // every return path restores the caller's current device
        int prev = 0;
        DeviceGuard() { cudaGetDevice(&prev); }
        ~DeviceGuard() { cudaSetDevice(prev); }
    } guard;
    CK(cudaSetDevice(device));  // the CK() returns below happen before any asynchronous work is queued
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

    size_t bpp = 0;
    for (int a = 0; a < narr; a++) bpp += (size_t)arrs[a].bytes_per_point;
    // HPGB_HOST_MIN_CHUNKS (default 1 = off): cut mid-sized calls (BASELINE config 1 is 10M points = 2.4 default chunks)
    // into at least that many pieces so that more of the first copy-in / last copy-out overlaps.  Measured and NOT
    // adopted: 10M pageable points take 7.1 ms as three chunks and 8.2 / 10.4 ms as eight (a chunk's fixed host cost --
    // three pool jobs, four CUDA calls, an event wait -- outweighs the overlap), 1e8 points are unaffected.
    int64_t cpb = chunk_points(bpp);
    {
        static const int64_t min_chunks = [] {
            const char *e = getenv("HPGB_HOST_MIN_CHUNKS");
            const int64_t v = e ? atoll(e) : 1;
            return v < 1 ? (int64_t)1 : v;
        }();
        const int64_t want = ((n + min_chunks - 1) / min_chunks + 1) & ~(int64_t)1;
        const int64_t floor_pts = (int64_t)1 << 18;
        if (want < cpb) cpb = want > floor_pts ? want : (floor_pts < cpb ? floor_pts : cpb);
    }
    const int64_t cp = cpb < n ? cpb : ((n + 1) & ~(int64_t)1);
    const int64_t nchunks = (n + cp - 1) / cp;
    const int nslots = nchunks < HPGB_PIPE_SLOTS ? (int)nchunks : HPGB_PIPE_SLOTS;  // no more slots than chunks
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
    for (int s = 0; s < nslots; s++) {
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
#define CKB(call)                                  \
    {                                              \
        cudaError_t e_ = (call);                   \
        if (e_ != cudaSuccess) {                   \
            rc = HPGB_E_CUDA + (int)e_;            \
            break;                                 \
        }                                          \
    }
    // every exit from here on goes through the common tail (synchronise the slot streams, restore the device):
    // an early return would leave kernels / copies in flight on the caller's buffers
    for (int64_t base = 0; base < n && rc == HPGB_OK; base += cp, slot = (slot + 1) % nslots) {
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
                cudaError_t e_ = cudaMemcpyAsync(dptr[a], src, bytes, cudaMemcpyHostToDevice, st);
                if (e_ != cudaSuccess) {
                    rc = HPGB_E_CUDA + (int)e_;
                    break;
                }
            }
        }
        if (rc != HPGB_OK) break;
        rc = launch(dptr, base, cn, c.d_status, st);
        if (rc != HPGB_OK) break;
        for (int a = 0; a < narr && rc == HPGB_OK; a++) {
            if (arrs[a].h_out) {
                char *dst = stage[a] ? c.hbuf[slot] + off[a] : arrs[a].h_out + (size_t)base * arrs[a].bytes_per_point;
                cudaError_t e_ = cudaMemcpyAsync(dst, dptr[a], (size_t)cn * arrs[a].bytes_per_point, cudaMemcpyDeviceToHost, st);
                if (e_ != cudaSuccess) rc = HPGB_E_CUDA + (int)e_;
            }
        }
        if (rc != HPGB_OK) break;
        if (any_stage) {
            CKB(cudaEventRecord(c.done[slot], st));
            slot_base[slot] = base;
            slot_n[slot] = cn;
        }
    }
#undef CKB
    if (any_stage && rc == HPGB_OK) {
        for (int k = 0; k < nslots && rc == HPGB_OK; k++) rc = drain((slot + k) % nslots);
    }
    for (int s = 0; s < HPGB_PIPE_SLOTS; s++) {
        cudaError_t e = cudaStreamSynchronize(c.stream[s]);
        if (e != cudaSuccess && rc == HPGB_OK) rc = HPGB_E_CUDA + (int)e;
    }
    if (rc == HPGB_OK) {
        cudaError_t e = cudaMemcpy(c.h_status, c.d_status, sizeof(hpgb_status_t), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess)
            rc = HPGB_E_CUDA + (int)e;
        else if (first_bad && c.h_status->first_bad != HPGB_NO_BAD_ELEMENT)
            *first_bad = (int64_t)c.h_status->first_bad;
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------
// One host array over several GPUs from ONE process (device == HPGB_DEVICE_ALL): the flat index range is cut
// into contiguous shards exactly like the reference cuts it over its worker threads (hpgeom/hpgeom.c:272-279:
// chunk = n / T, the first n % T workers take one extra element), one host thread per GPU runs the ordinary
// single-device pipeline on its shard, results land in the caller's output arrays.  No collective: shards are
// independent.  Element errors keep their meaning: every shard reports GLOBAL indices, the smallest wins.
// ---------------------------------------------------------------------------------------------
namespace {
std::mutex g_devset_mutex;
std::vector<int> g_devset;  // empty = every visible device
const int64_t kMinShardPoints = (int64_t)1 << 18;  // below this a second GPU costs more (launch + copy latency) than it brings
}  // namespace

extern "C" int hpgb_host_set_devices(const int *devices, int n_dev) {
    const int ndev = hpgb_device_count();
    if (n_dev < 0 || n_dev > 64 || (n_dev > 0 && !devices)) return HPGB_E_BAD_ARGUMENT;
    for (int i = 0; i < n_dev; i++) {
        if (devices[i] < 0 || devices[i] >= ndev) return HPGB_E_BAD_ARGUMENT;
        for (int j = 0; j < i; j++)
            if (devices[j] == devices[i]) return HPGB_E_BAD_ARGUMENT;
    }
    std::lock_guard<std::mutex> lock(g_devset_mutex);
    g_devset.assign(devices, devices + n_dev);
    return HPGB_OK;
}

extern "C" int hpgb_host_get_devices(int *devices, int cap) {
    std::vector<int> set;
    {
        std::lock_guard<std::mutex> lock(g_devset_mutex);
        set = g_devset;
    }
    if (set.empty())
        for (int d = 0; d < hpgb_device_count() && d < 64; d++) set.push_back(d);
    for (int i = 0; i < (int)set.size() && i < cap; i++) devices[i] = set[i];
    return (int)set.size();
}

int hpgb_host_first_device(int device) {
    if (device != HPGB_DEVICE_ALL) return device;
    int d[64];
    return hpgb_host_get_devices(d, 64) > 0 ? d[0] : 0;
}

int hpgb_host_pipeline(int device, int64_t n, const hpgb_pipe_arr *arrs, int narr, const hpgb_pipe_launch &launch,
                       int64_t *first_bad) {
    if (device != HPGB_DEVICE_ALL) return pipeline_one(device, n, arrs, narr, launch, first_bad);
    if (first_bad) *first_bad = -1;
    if (n <= 0) return HPGB_OK;
    int devs[64];
    int G = hpgb_host_get_devices(devs, 64);
    if (G <= 0) return HPGB_E_NO_DEVICE;
    if ((int64_t)G > (n + kMinShardPoints - 1) / kMinShardPoints) G = (int)((n + kMinShardPoints - 1) / kMinShardPoints);
    if (G <= 1) return pipeline_one(devs[0], n, arrs, narr, launch, first_bad);
    std::vector<int> rc(G, HPGB_OK);
    std::vector<int64_t> fb(G, -1);
    const int64_t chunk = n / G, rem = n % G;
    auto shard = [&](int r) {
        const int64_t lo = r * chunk + (r < rem ? r : rem), cn = chunk + (r < rem ? 1 : 0);
        hpgb_pipe_arr sub[HPGB_PIPE_MAX_ARR];
        for (int a = 0; a < narr; a++) {
            sub[a] = arrs[a];
            if (sub[a].h_in) sub[a].h_in += (size_t)lo * (size_t)arrs[a].bytes_per_point;
            if (sub[a].h_out) sub[a].h_out += (size_t)lo * (size_t)arrs[a].bytes_per_point;
        }
        rc[r] = pipeline_one(devs[r], cn, sub, narr,
                             [&launch, lo](void **d, int64_t base, int64_t c, hpgb_status_t *st, cudaStream_t s) {
                                 return launch(d, lo + base, c, st, s);
                             },
                             &fb[r]);
    };
    std::vector<std::thread> workers;
    for (int r = 1; r < G; r++) workers.emplace_back(shard, r);
    shard(0);
    for (auto &w : workers) w.join();
    int out_rc = HPGB_OK;
    int64_t best = -1;
    for (int r = 0; r < G; r++) {
        if (rc[r] != HPGB_OK && out_rc == HPGB_OK) out_rc = rc[r];
        if (fb[r] >= 0 && (best < 0 || fb[r] < best)) best = fb[r];
    }
    if (first_bad) *first_bad = best;
    return out_rc;
}

extern "C" int hpgb_host_release(void) {
    int prev = 0;
    cudaGetDevice(&prev);
    for (int d = 0; d < 64; d++) {
        std::lock_guard<std::mutex> lock(g_ctx_mutex[d]);
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

// The staging copy of the host pipeline on its own (the pool's parallel memcpy): exported so that the CPU test tier can
// check it without a GPU and tools/host_copy_rate.py can time it.  (Measured and rejected: non-temporal SSE2 stores for
// the shares -- 44.6 against 55.3 GB/s for glibc's memcpy on 8 threads.)
extern "C" int hpgb_host_copy(void *dst, const void *src, int64_t bytes) {
    if (bytes < 0 || (bytes > 0 && (!dst || !src))) return HPGB_E_BAD_ARGUMENT;
    if (bytes) CopyPool::get().copy((char *)dst, (const char *)src, (size_t)bytes);
    return HPGB_OK;
}
