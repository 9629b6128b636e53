// Shared helpers for libpgb200 (sm_100a).  Internal header — the public ABI is include/pgb200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "pgb200.h"

namespace pgb {

// thread-local last-error string, retrieved through pgb_last_error()
char *error_buffer();
int fail(const char *fmt, ...);

#define PGB_CUDA_OK(expr)                                                                              \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return ::pgb::fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define PGB_LAUNCH_OK(name)                                                                            \
    do {                                                                                               \
        cudaError_t _e = cudaGetLastError();                                                           \
        if (_e != cudaSuccess)                                                                         \
            return ::pgb::fail("launch of %s failed: %s", name, cudaGetErrorString(_e));              \
    } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid size for simple grid-stride kernels: enough CTAs to fill the 148 SMs a few times over
inline int stride_grid(int64_t n, int block) {
    int64_t want = ceil_div(n, block);
    const int64_t cap = 148 * 16;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

int sm_count();          // SMs of the CURRENT device
int current_device();    // cudaGetDevice, clamped to [0, PGB_MAX_DEVICES)
constexpr int PGB_MAX_DEVICES = 64;
// Function attributes (opt-in shared memory, carveout) and occupancy are per device: launchers keep one cache
// slot per device instead of a process-wide static, so a process that drives several GPUs configures each.
struct PerDeviceInt {
    int v[PGB_MAX_DEVICES] = {};
    int &here() { return v[current_device()]; }
};

// ---- device helpers ------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of one double per thread; result valid in thread 0.  `scratch` needs 32 doubles.
__device__ __forceinline__ double block_sum(double v, double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    const int nwarps = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nwarps) ? scratch[threadIdx.x] : 0.0;
    if (warp == 0) v = warp_sum(v);
    return v;
}

// block-wide maximum of one non-negative double per thread; result valid in thread 0
__device__ __forceinline__ double block_max(double v, double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    const int nwarps = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nwarps) ? scratch[threadIdx.x] : 0.0;
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    return v;
}

// streaming (evict-first) loads for data touched once per iteration: CSR indices / weights
__device__ __forceinline__ int ld_stream(const int *p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream(const float *p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace pgb
