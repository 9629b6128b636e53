// Microbenchmark (not product code): cost of flushing hsell pieces as coalesced 128-byte RED.ADD.F32 into a
// y vector (L2 atomics) versus plain 128-byte stores into partial rows.  148 CTAs x 32 warps, each warp flushes
// rows at pseudo-random (slice) positions of a 64 MB array, with a skew option (hot rows, like hub slices).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t mixu(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
template <int MODE, typename T>   // 0: store, 1: red.add, 2: red.add with 1/8 of the rows hot (4096 rows)
__global__ void __launch_bounds__(1024, 1) flush_kernel(T *y, int n_rows, int per_warp, T *out) {
    const int lane = threadIdx.x & 31;
    const uint32_t wid = blockIdx.x * 32 + (threadIdx.x >> 5);
    T v = (T)(lane + 1);
    for (int i = 0; i < per_warp; ++i) {
        uint32_t h = mixu(wid * 65537u + i * 7919u + 11u);
        uint32_t row = h % (uint32_t)n_rows;
        if (MODE == 2 && (h >> 29) == 0) row = (h >> 8) & 4095u;
        T *p = y + (size_t)row * 32 + lane;
        if (MODE == 0) *p = v;
        else atomicAdd(p, v);
        v += (T)1;
    }
    if (v == (T)-1) out[0] = v;
}
template <typename T>
void run_all(const char *tname) {
    const int n_rows = (64 << 20) / (32 * sizeof(T));
    T *y, *out;
    CK(cudaMalloc(&y, (size_t)n_rows * 32 * sizeof(T)));
    CK(cudaMalloc(&out, 64));
    CK(cudaMemset(y, 0, (size_t)n_rows * 32 * sizeof(T)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int per_warp = 450;   // 148*32*450 = 2.13 M rows per launch (~ pieces of RMAT-24)
    auto t = [&](auto kern, const char *name) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(e0));
            kern<<<148, 1024>>>(y, n_rows, per_warp, out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
        }
        printf("%s %-40s %7.3f ms for %.2f M rows  (%.1f cycles per row per SM)\n", tname, name, best, 148 * 32 * per_warp / 1e6,
               best * 1e-3 * 1.965e9 / (32.0 * per_warp));
    };
    t(flush_kernel<0, T>, "128 B stores to random rows");
    t(flush_kernel<1, T>, "128 B RED.ADD to random rows");
    t(flush_kernel<2, T>, "128 B RED.ADD, 1/8 into 4096 hot rows");
    CK(cudaFree(y));
}
int main() {
    run_all<float>("f32");
    run_all<double>("f64");
    return 0;
}
