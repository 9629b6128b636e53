// Microbenchmark (not product code): shared-memory gather rate per SM on B200 for the address patterns the hsell
// hub path can produce.  1024 threads per CTA, one CTA per SM, 128 KB table; every warp issues ROUNDS x 2 LDS.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/lds_rate.cu -o lds_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mixu(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// pattern 0: broadcast (all lanes one address), 1: consecutive (lane i -> base + i), 2: conflict-free random rows
// (bank = lane), 3: fully random, 4: random with the bank-aware rotation (bank = (lane + k) % 32, random row)
template <int UNROLL, typename V>
__global__ void __launch_bounds__(1024, 1) lds_kernel(int pattern, int rounds, float *out, int H) {
    extern __shared__ float s_z[];
    for (int i = threadIdx.x; i < H; i += 1024) s_z[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t st = mixu(threadIdx.x * 977u + blockIdx.x * 31337u + 1u);
    float a0 = 0.f, a1 = 0.f;
    const uint32_t HV = H / (sizeof(V) / 4);   // a power of two
    // indices are prepared once (the product kernel gets them from the index stream); every round advances each
    // by a multiple of 32 elements, which keeps its bank
    uint32_t idx0[UNROLL], idx1[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
        st = st * 1664525u + 1013904223u;
        const uint32_t h0 = st >> 8, h1 = (st * 2654435761u) >> 8;
        uint32_t i0, i1;
        if (pattern == 0) { i0 = k; i1 = k + 7; }
        else if (pattern == 1) { i0 = (k * 32 + lane) % HV; i1 = (k * 32 + 4096 + lane) % HV; }
        else if (pattern == 2) { i0 = ((h0 % (HV / 32)) * 32) | lane; i1 = ((h1 % (HV / 32)) * 32) | lane; }
        else if (pattern == 3) { i0 = h0 % HV; i1 = h1 % HV; }
        else { i0 = ((h0 % (HV / 32)) * 32) | ((lane + k) & 31); i1 = ((h1 % (HV / 32)) * 32) | ((lane + k + 16) & 31); }
        idx0[k] = i0;
        idx1[k] = i1;
    }
    for (int r = 0; r < rounds; r += UNROLL) {
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {
            const V v0 = reinterpret_cast<const V *>(s_z)[idx0[k]];
            const V v1 = reinterpret_cast<const V *>(s_z)[idx1[k]];
            a0 += *reinterpret_cast<const float *>(&v0);
            a1 += *reinterpret_cast<const float *>(&v1);
            idx0[k] = (idx0[k] + 32u * 37u) & (HV - 1);
            idx1[k] = (idx1[k] + 32u * 101u) & (HV - 1);
        }
    }
    if (a0 + a1 == 123.456f) out[0] = a0;
}

int main() {
    const int H = 32768;
    float *out;
    CK(cudaMalloc(&out, 64));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int rounds = 4096;
    auto time_it = [&](auto kern, const char *name, int pattern, int threads) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 132 * 1024));
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            kern<<<148, threads, (H + 32) * 4>>>(pattern, rounds, out, H);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        const double cyc = best * 1e-3 * 1.965e9;
        const double lds_per_sm = (double)rounds * 2 * (threads / 32);
        printf("%-34s pattern %d  warps %2d  %7.3f ms  %5.2f cyc per LDS per SM\n", name, pattern, threads / 32, best, cyc / lds_per_sm);
    };
    const char *names[] = {"broadcast", "consecutive", "conflict-free random rows", "fully random", "bank-rotated random rows"};
    for (int p = 0; p < 5; ++p) {
        for (int threads : {1024, 512, 256}) time_it(lds_kernel<8, float>, names[p], p, threads);
    }
    for (int p = 1; p < 4; ++p) time_it(lds_kernel<8, float2>, "LDS.64", p, 1024);
    for (int p = 1; p < 4; ++p) time_it(lds_kernel<16, float>, "unroll 16", p, 1024);
    return 0;
}
