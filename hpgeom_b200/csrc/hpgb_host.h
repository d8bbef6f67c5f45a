// hpgb_host.h -- internal interface between the C-ABI entry points (hpgb_kernels.cu) and the
// chunked host pipeline (hpgb_host.cu).
#ifndef HPGB_HOST_H
#define HPGB_HOST_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>

#include "../../include/hpgb.h"

#define HPGB_PIPE_SLOTS 3
#define HPGB_PIPE_MAX_ARR 8

// one operand of a host-pointer call: h_in != NULL -> copied to the device before the kernel,
// h_out != NULL -> copied back after it; bytes_per_point = bytes this operand holds per point
struct hpgb_pipe_arr {
    const char *h_in;
    char *h_out;
    int64_t bytes_per_point;
};

// launch(device pointers in operand order, global index of the chunk's first point, points in
// the chunk, device status, stream) -> HPGB_* code
typedef std::function<int(void **, int64_t, int64_t, hpgb_status_t *, cudaStream_t)> hpgb_pipe_launch;

// device scratch of the variable-length entry points (range tables, search frontiers): stream-ordered
// allocations on the legacy stream from the device's default memory pool, whose release threshold is
// raised once so that freed blocks are kept for the next call (a cudaMalloc / cudaFree pair costs more
// than a small query)
cudaError_t hpgb_dev_alloc(void **p, size_t bytes);
void hpgb_dev_free(void *p);

// device >= 0: that GPU; device == HPGB_DEVICE_ALL: contiguous shards over the device set (hpgb_host_set_devices),
// one host thread per GPU
int hpgb_host_pipeline(int device, int64_t n, const hpgb_pipe_arr *arrs, int narr,
                       const hpgb_pipe_launch &launch, int64_t *first_bad);
// entry points whose device scratch lives on ONE GPU (range tables, query plans): HPGB_DEVICE_ALL -> first of the set
int hpgb_host_first_device(int device);

#endif
