// common.cuh - shared host/device helpers for libhdp_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "hdp_b200.h"

namespace hdp {

// Number of kernels this library has launched (bench.py reports it as gpu_launches).
extern int64_t g_launch_count;

inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? HDP_B200_OK : (int)e; }

#define HDP_CUDA_TRY(expr)                                   \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) return (int)_e;               \
    } while (0)

// After every launch: count it and surface launch-configuration errors without synchronising.
#define HDP_LAUNCH_CHECK()                                   \
    do {                                                     \
        ::hdp::g_launch_count++;                             \
        cudaError_t _e = cudaGetLastError();                 \
        if (_e != cudaSuccess) return (int)_e;               \
    } while (0)

// Optional per-kernel CUDA-event timing (hdp_b200_timing_enable / hdp_b200_timing_read).
enum KernelId { kNormalize = 1, kThrGeneric = 2, kHotWords = 3, kScan = 4, kUnpackMask = 5, kThrSeg = 6, kThrRanked = 7, kMeasure = 8 };
struct KernelTimer {
    bool on;
    cudaStream_t st;
    int slot;
    KernelTimer(int id, cudaStream_t st);
    ~KernelTimer();
};

constexpr size_t kAlign = 256;
inline size_t align_up(size_t v, size_t a = kAlign) { return (v + a - 1) / a * a; }

// Bump allocator over the caller-provided workspace.
struct Carver {
    char *base;
    size_t size;
    size_t off = 0;
    Carver(void *p, size_t n) : base((char *)p), size(n) {}
    template <typename T>
    T *take(size_t count) {
        size_t bytes = align_up(count * sizeof(T));
        T *r = (T *)(base ? base + off : nullptr);
        off += bytes;
        return r;
    }
    bool ok() const { return off <= size; }
};

// Copies the strided measure array into a time-major, cell-contiguous [T, C] buffer.
int normalize_layout(const float *src, int64_t C, int64_t T, int64_t ld_t, int64_t ld_c, float *dst, cudaStream_t st);

}  // namespace hdp
