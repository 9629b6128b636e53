// Microbenchmark (not product code): ceilings for 4-byte random gathers on B200.
//   A: global gathers (indices streamed coalesced), table sizes 64 MB / 8 MB / 128 KB, uniform or skewed
//   B: shared-memory gathers from a 48K-float table
//   C: mixed — indices below H served from shared memory, the rest from global
// Prints giga-gathers per second for each variant.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// RMAT-like popularity: each of `bits` bits is 1 with probability 0.24; then map to "degree order":
// popcount-sorted rank approximated by bit-reversal-free trick: we simply keep the raw id (hubs = few one-bits)
__global__ void make_idx(int* idx, int64_t m, int bits, int mode, int table) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t h = mix64(i * 0xD1342543DE82EF95ull + 12345);
        uint32_t v = 0;
        if (mode == 0) {
            v = (uint32_t)(h % (uint64_t)table);
        } else {
            uint64_t hh = h;
            for (int b = 0; b < bits; ++b) {
                if ((b & 3) == 0) hh = mix64(hh);
                uint32_t u = (uint32_t)(hh >> ((b & 3) * 16)) & 0xFFFF;
                v = (v << 1) | (u < 15729 ? 1u : 0u);  // 0.24 * 65536
            }
        }
        idx[i] = (int)v;
    }
}

template <int IPT, typename T>
__global__ void __launch_bounds__(256, 4) gather_global(const int* __restrict__ idx, const T* __restrict__ z, int64_t m, T* out) {
    T acc = 0;
    const int64_t tile = 256 * IPT;
    for (int64_t base = blockIdx.x * tile; base < m; base += (int64_t)gridDim.x * tile) {
        int c[IPT];
#pragma unroll
        for (int s = 0; s < IPT; ++s) { int64_t i = base + s * 256 + threadIdx.x; c[s] = i < m ? __ldcs(idx + i) : 0; }
#pragma unroll
        for (int s = 0; s < IPT; ++s) acc += __ldg(z + c[s]);
    }
    if (acc == (T)123456789) out[0] = acc;
}

template <int IPT>
__global__ void __launch_bounds__(1024, 1) gather_mixed(const int* __restrict__ idx, const float* __restrict__ z, int64_t m, int H, float* out) {
    extern __shared__ float hub[];
    for (int i = threadIdx.x; i < H; i += blockDim.x) hub[i] = z[i];
    __syncthreads();
    float acc = 0;
    const int64_t tile = 1024 * IPT;
    for (int64_t base = blockIdx.x * tile; base < m; base += (int64_t)gridDim.x * tile) {
        int c[IPT];
#pragma unroll
        for (int s = 0; s < IPT; ++s) { int64_t i = base + s * 1024 + threadIdx.x; c[s] = i < m ? __ldcs(idx + i) : 0; }
#pragma unroll
        for (int s = 0; s < IPT; ++s) acc += (c[s] < H) ? hub[c[s]] : __ldg(z + c[s]);
    }
    if (acc == 123456789.f) out[0] = acc;
}

// popcount-rank relabel: id -> rank in (popcount, value) order, so hubs (few one-bits) get small ids
__global__ void relabel(int* idx, int64_t m, const int* rank) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) idx[i] = rank[idx[i]];
}

static float time_it(void (*launch)(void*), void* ctx, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(ctx); launch(ctx);
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) launch(ctx);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

struct Ctx { const int* idx; const float* z; const double* zd; int64_t m; float* out; double* outd; int H; int grid; };

int main() {
    const int64_t m = 512ll << 20;
    const int bits = 24, n = 1 << bits;
    int* idx; float* z; double* zd; float* out; double* outd; int* rank;
    cudaMalloc(&idx, m * 4); cudaMalloc(&z, (size_t)n * 4); cudaMalloc(&zd, (size_t)n * 8); cudaMalloc(&out, 64); cudaMalloc(&outd, 64);
    cudaMemset(z, 0, (size_t)n * 4); cudaMemset(zd, 0, (size_t)n * 8);
    // host-side popcount rank
    int* hrank = (int*)malloc((size_t)n * 4);
    {
        int64_t* cnt = (int64_t*)calloc(bits + 2, 8);
        for (int i = 0; i < n; ++i) cnt[__builtin_popcount(i) + 1]++;
        for (int k = 1; k <= bits + 1; ++k) cnt[k] += cnt[k - 1];
        for (int i = 0; i < n; ++i) hrank[i] = (int)cnt[__builtin_popcount(i)]++;
        free(cnt);
    }
    cudaMalloc(&rank, (size_t)n * 4); cudaMemcpy(rank, hrank, (size_t)n * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(gather_mixed<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    Ctx c{idx, z, zd, m, out, outd, 0, 148 * 4};
    auto report = [&](const char* name, float ms) { printf("%-58s %8.3f ms  %8.1f Ggather/s\n", name, ms, m / (ms * 1e-3) / 1e9); fflush(stdout); };

    make_idx<<<1184, 256>>>(idx, m, bits, 0, n);
    report("A global f32 uniform over 64MB (IPT 9, 4 CTA/SM)", time_it([](void* p) { Ctx* c = (Ctx*)p; gather_global<9, float><<<c->grid, 256>>>(c->idx, c->z, c->m, c->out); }, &c, 5));
    report("A global f64 uniform over 128MB", time_it([](void* p) { Ctx* c = (Ctx*)p; gather_global<9, double><<<c->grid, 256>>>(c->idx, c->zd, c->m, c->outd); }, &c, 5));
    report("A global f32 uniform, IPT 16", time_it([](void* p) { Ctx* c = (Ctx*)p; gather_global<16, float><<<c->grid, 256>>>(c->idx, c->z, c->m, c->out); }, &c, 5));
    make_idx<<<1184, 256>>>(idx, m, bits, 0, 2 << 20);
    report("A global f32 uniform over 8MB", time_it([](void* p) { Ctx* c = (Ctx*)p; gather_global<9, float><<<c->grid, 256>>>(c->idx, c->z, c->m, c->out); }, &c, 5));
    make_idx<<<1184, 256>>>(idx, m, bits, 0, 24 << 10);
    report("A global f32 uniform over 96KB (L1 resident)", time_it([](void* p) { Ctx* c = (Ctx*)p; gather_global<9, float><<<c->grid, 256>>>(c->idx, c->z, c->m, c->out); }, &c, 5));
    c.H = 48 << 10; c.grid = 148;
    make_idx<<<1184, 256>>>(idx, m, bits, 0, 48 << 10);
    report("B shared f32 uniform over 192KB hub table (1024 thr)", time_it([](void* p) { Ctx* c = (Ctx*)p; gather_mixed<9><<<c->grid, 1024, c->H * 4>>>(c->idx, c->z, c->m, c->H, c->out); }, &c, 5));
    // RMAT popularity
    make_idx<<<1184, 256>>>(idx, m, bits, 1, n);
    c.grid = 148 * 4;
    report("A global f32 RMAT popularity, raw ids", time_it([](void* p) { Ctx* c = (Ctx*)p; gather_global<9, float><<<c->grid, 256>>>(c->idx, c->z, c->m, c->out); }, &c, 5));
    relabel<<<1184, 256>>>(idx, m, rank);
    report("A global f32 RMAT popularity, degree-ranked ids", time_it([](void* p) { Ctx* c = (Ctx*)p; gather_global<9, float><<<c->grid, 256>>>(c->idx, c->z, c->m, c->out); }, &c, 5));
    c.grid = 148;
    for (int H : {12 << 10, 24 << 10, 48 << 10, 55 << 10}) {
        c.H = H;
        char name[128]; snprintf(name, sizeof(name), "C mixed RMAT popularity ranked, hub table %dK floats", H >> 10);
        report(name, time_it([](void* p) { Ctx* c = (Ctx*)p; gather_mixed<9><<<c->grid, 1024, c->H * 4>>>(c->idx, c->z, c->m, c->H, c->out); }, &c, 5));
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
