// Microbenchmark (not product code): per-SM cost of the three gather paths the hsell kernel can use on B200,
// alone and running together in one persistent CTA of 32 warps (the shape of hsell_gather_kernel):
//   lds  : hub path   — stream 32-bit words (two 16-bit columns) and gather twice from a 128 KB shared block
//   lsu  : tail path  — stream 32-bit columns and gather z[col] with LDG (L1TEX LSU pipe)
//   tex  : tail path  — the same gathers with tex1Dfetch (L1TEX TEX pipe)
// Work is expressed in ROUNDS (one word per lane); every warp takes chunks of 32 rounds from a queue.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/gather_paths.cu -o gather_paths
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void fill_words(uint32_t *w, int64_t m, uint32_t H, int conflict_free) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t h = mix64(i);
        uint32_t a = (uint32_t)(h % H), b = (uint32_t)((h >> 32) % H);
        if (conflict_free) {   // lane l reads bank l: the floor of the shared-memory path
            const uint32_t lane = (uint32_t)(i & 31);
            a = (a & ~31u) | lane;
            b = (b & ~31u) | lane;
        }
        w[i] = a | (b << 16);
    }
}
__global__ void fill_cols(int32_t *c, int64_t m, uint32_t lo, uint32_t hi) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
        c[i] = (int32_t)(lo + mix64(i * 77 + 5) % (hi - lo));
}

struct P {
    const uint32_t *words;   // hub stream, [round][lane]
    const int32_t *cols;     // tail stream
    const float *z;
    cudaTextureObject_t tex;
    int hub_chunks, tail_chunks;
    unsigned *queues;        // [0] hub, [1] tail
    float *out;
    int tail_warps;          // warps per CTA that take tail chunks first
    int tail_mode;           // 0 lsu, 1 tex, 2 lsu with one 128-bit index load per 4 rounds
    int hub_mode;            // 0 LDG.32 words, 1 LDG.128 words ([round/4][lane][4] layout)
    int H;
    int pf_dist;             // L2 prefetch distance in hub chunks (0 = off)
    int wrap;                // > 0: hub chunk ids wrap at this count (stream stays in L2)
    int pf_mode;             // 0: cp.async.bulk.prefetch.L2 (one lane, 4 KB), 1: prefetch.global.L2 per lane (32 x 128 B)
};

__device__ __forceinline__ float hub_chunk(const P &p, int u, const float *s_z, int lane) {
    float a0 = 0.f, a1 = 0.f;
    if (p.wrap > 0) u = u % p.wrap;   // L2-resident variant: the same few chunks over and over
    if (p.hub_mode == 2) {            // stream only: no shared-memory gathers
        const uint32_t *d = p.words + (int64_t)u * 1024 + lane;
        uint32_t w[8], nx[8];
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = __ldcs(d + k * 32);
#pragma unroll
        for (int bt = 0; bt < 4; ++bt) {
            if (bt < 3) {
#pragma unroll
                for (int k = 0; k < 8; ++k) nx[k] = __ldcs(d + ((bt + 1) * 8 + k) * 32);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) s += w[k];
#pragma unroll
            for (int k = 0; k < 8; ++k) w[k] = nx[k];
        }
        return __uint_as_float(s);
    }
    if (p.hub_mode == 3) {            // gathers only: one word per lane reused for the 32 rounds (no stream)
        uint32_t w = p.words[(int64_t)u * 1024 + lane];
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            a0 += s_z[w & 0xffffu];
            a1 += s_z[w >> 16];
            w = ((w + 32u) & 0x7fff7fffu);
        }
        return a0 + a1;
    }
    if (p.hub_mode == 0) {
        const uint32_t *d = p.words + (int64_t)u * 1024 + lane;
        uint32_t w[8], nx[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = __ldcs(d + k * 32);
#pragma unroll
        for (int bt = 0; bt < 4; ++bt) {
            if (bt < 3) {
#pragma unroll
                for (int k = 0; k < 8; ++k) nx[k] = __ldcs(d + ((bt + 1) * 8 + k) * 32);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                a0 += s_z[w[k] & 0xffffu];
                a1 += s_z[w[k] >> 16];
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) w[k] = nx[k];
        }
    } else {
        const uint4 *d = reinterpret_cast<const uint4 *>(p.words + (int64_t)u * 1024) + lane;
        uint4 w[2], nx[2];
        w[0] = __ldcs(d);
        w[1] = __ldcs(d + 32);
#pragma unroll
        for (int bt = 0; bt < 4; ++bt) {
            if (bt < 3) {
                nx[0] = __ldcs(d + ((bt + 1) * 2) * 32);
                nx[1] = __ldcs(d + ((bt + 1) * 2 + 1) * 32);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                a0 += s_z[w[k].x & 0xffffu]; a1 += s_z[w[k].x >> 16];
                a0 += s_z[w[k].y & 0xffffu]; a1 += s_z[w[k].y >> 16];
                a0 += s_z[w[k].z & 0xffffu]; a1 += s_z[w[k].z >> 16];
                a0 += s_z[w[k].w & 0xffffu]; a1 += s_z[w[k].w >> 16];
            }
            w[0] = nx[0];
            w[1] = nx[1];
        }
    }
    return a0 + a1;
}

__device__ __forceinline__ float tail_chunk(const P &p, int u, int lane) {
    float a0 = 0.f, a1 = 0.f;
    if (p.tail_mode == 2) {
        const int4 *d = reinterpret_cast<const int4 *>(p.cols + (int64_t)u * 1024) + lane;
        int4 c[2], nx[2];
        c[0] = __ldcs(d);
        c[1] = __ldcs(d + 32);
#pragma unroll
        for (int bt = 0; bt < 4; ++bt) {
            float x[8];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                x[4 * k] = __ldg(p.z + c[k].x); x[4 * k + 1] = __ldg(p.z + c[k].y);
                x[4 * k + 2] = __ldg(p.z + c[k].z); x[4 * k + 3] = __ldg(p.z + c[k].w);
            }
            if (bt < 3) {
                nx[0] = __ldcs(d + ((bt + 1) * 2) * 32);
                nx[1] = __ldcs(d + ((bt + 1) * 2 + 1) * 32);
            }
#pragma unroll
            for (int k = 0; k < 8; k += 2) { a0 += x[k]; a1 += x[k + 1]; }
            c[0] = nx[0];
            c[1] = nx[1];
        }
        return a0 + a1;
    }
    const int32_t *d = p.cols + (int64_t)u * 1024 + lane;
    int32_t c[8], nx[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) c[k] = __ldcs(d + k * 32);
#pragma unroll
    for (int bt = 0; bt < 4; ++bt) {
        float x[8];
        if (p.tail_mode == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = tex1Dfetch<float>(p.tex, c[k]);
        } else if (p.tail_mode == 3) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(x[k]) : "l"(p.z + c[k]));
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = __ldg(p.z + c[k]);
        }
        if (bt < 3) {
#pragma unroll
            for (int k = 0; k < 8; ++k) nx[k] = __ldcs(d + ((bt + 1) * 8 + k) * 32);
        }
#pragma unroll
        for (int k = 0; k < 8; k += 2) { a0 += x[k]; a1 += x[k + 1]; }
#pragma unroll
        for (int k = 0; k < 8; ++k) c[k] = nx[k];
    }
    return a0 + a1;
}

__global__ void __launch_bounds__(1024, 1) paths_kernel(const P p) {
    extern __shared__ float s_z[];
    for (int i = threadIdx.x; i < p.H; i += 1024) s_z[i] = p.z[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    const bool tail_first = warp < p.tail_warps;
    // hub chunks: a static contiguous range per CTA behind a shared-memory counter (as in hsell_gather_kernel);
    // tail chunks: one global queue
    __shared__ int s_next;
    const int hub_lo = (int)((int64_t)p.hub_chunks * blockIdx.x / gridDim.x);
    const int hub_hi = (int)((int64_t)p.hub_chunks * (blockIdx.x + 1) / gridDim.x);
    if (threadIdx.x == 0) s_next = hub_lo;
    __syncthreads();
    for (int phase = 0; phase < 2; ++phase) {
        const int kind = (phase == 0) == tail_first ? 1 : 0;   // 1 tail, 0 hub
        const int limit = kind ? p.tail_chunks : hub_hi;
        int u = 0;
        if (lane == 0) u = kind ? (int)atomicAdd(p.queues + 1, 1u) : atomicAdd(&s_next, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        while (u < limit) {
            int nu = 0;
            if (lane == 0) {
                nu = kind ? (int)atomicAdd(p.queues + 1, 1u) : atomicAdd(&s_next, 1);
                if (p.pf_dist > 0) {
                    if (!kind && nu + p.pf_dist < hub_hi && p.pf_mode == 0)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.words + (int64_t)(nu + p.pf_dist) * 1024), "r"(4096) : "memory");
                    if (kind && nu < limit)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.cols + (int64_t)nu * 1024), "r"(4096) : "memory");
                }
            }
            if (p.pf_dist > 0 && p.pf_mode == 1 && !kind) {   // one 128-byte line per lane: the whole chunk pf_dist ahead
                const int t = u + p.pf_dist;
                if (t < hub_hi) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.words + (int64_t)t * 1024 + lane * 32));
            }
            acc += kind ? tail_chunk(p, u, lane) : hub_chunk(p, u, s_z, lane);
            u = __shfl_sync(0xffffffffu, nu, 0);
        }
    }
    if (acc == 123456.789f) p.out[0] = acc;
}

int main(int argc, char **argv) {
    const int64_t n = 1 << 24;
    const int H = 32768;
    const int hub_chunks = 320000, tail_chunks = 64000;   // ~ RMAT-24: 10 M hub rounds, 2 M tail rounds
    uint32_t *words;
    int32_t *cols;
    float *z, *out;
    unsigned *queues;
    CK(cudaMalloc(&words, (size_t)hub_chunks * 4096));
    CK(cudaMalloc(&cols, (size_t)tail_chunks * 4096));
    CK(cudaMalloc(&z, n * 4));
    CK(cudaMalloc(&out, 64));
    CK(cudaMalloc(&queues, 64));
    CK(cudaMemset(z, 0, n * 4));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = z;
    rd.res.linear.desc = cudaCreateChannelDesc<float>();
    rd.res.linear.sizeInBytes = n * 4;
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    CK(cudaFuncSetAttribute(paths_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 132 * 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    int pf = 0, wrap = 0, pf_mode = 0;
    auto run = [&](const char *name, int hc, int tc, int tail_warps, int tail_mode, int hub_mode) {
        P p{words, cols, z, tex, hc, tc, queues, out, tail_warps, tail_mode, hub_mode, H, pf, wrap, pf_mode};
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaMemsetAsync(queues, 0, 64));
            CK(cudaEventRecord(e0));
            paths_kernel<<<148, 1024, (H + 32) * 4>>>(p);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
        }
        const double cyc = best * 1e-3 * 1.965e9;
        const double rounds_sm = ((double)hc + tc) * 32 / 148;
        printf("%-44s %7.3f ms  %6.2f cyc/round/SM  (hub %d tail %d chunks)\n", name, best, cyc / rounds_sm, hc, tc);
    };
    if (argc > 1 && argv[1][0] == 'l') {   // how much L1 do the tail gathers need?  (dynamic shared memory sweeps the carve-out)
        fill_cols<<<1024, 256>>>(cols, (int64_t)tail_chunks * 1024, 1 << 21, 1 << 24);
        CK(cudaDeviceSynchronize());
        CK(cudaFuncSetAttribute(paths_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        for (int kb : {129, 160, 192, 224}) {
            printf("---- %d KB of dynamic shared memory per CTA\n", kb);
            for (int mode : {0, 3, 1}) {
                P p{words, cols, z, tex, 0, tail_chunks, queues, out, 32, mode, 0, H, 0, 0, 0};
                float best = 1e9f;
                for (int rep = 0; rep < 4; ++rep) {
                    CK(cudaMemsetAsync(queues, 0, 64));
                    CK(cudaEventRecord(e0));
                    paths_kernel<<<148, 1024, kb * 1024>>>(p);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (rep && ms < best) best = ms;
                }
                printf("tail only, %-28s %7.3f ms\n", mode == 0 ? "LDG (__ldg)" : mode == 3 ? "LDG L1::no_allocate" : "TEX", best);
            }
        }
        return 0;
    }
    if (argc > 1 && argv[1][0] == 'p') {   // does an L2 prefetch shorten the stream's latency?
        fill_words<<<1024, 256>>>(words, (int64_t)hub_chunks * 1024, H, 0);
        CK(cudaDeviceSynchronize());
        for (int mode = 0; mode < 2; ++mode)
            for (int d : {0, 8, 32, 64, 128}) {
                pf = d;
                pf_mode = mode;
                printf("---- pf_mode %d (%s) distance %d chunks\n", mode, mode ? "prefetch.global.L2 per lane" : "cp.async.bulk.prefetch.L2", d);
                run("hub: stream only", hub_chunks, 0, 0, 0, 2);
                run("hub: stream + 2 LDS per round", hub_chunks, 0, 0, 0, 0);
            }
        return 0;
    }
    if (argc > 1) {   // decomposition of the hub path
        for (int cf = 0; cf < 2; ++cf) {
            fill_words<<<1024, 256>>>(words, (int64_t)hub_chunks * 1024, H, cf);
            CK(cudaDeviceSynchronize());
            for (int w = 0; w < 2; ++w) {
                wrap = w ? 4096 : 0;
                printf("---- %s banks, words %s\n", cf ? "conflict-free" : "random", w ? "L2-resident (16 MB window)" : "from HBM");
                run("hub: stream + 2 LDS per round", hub_chunks, 0, 0, 0, 0);
                run("hub: stream only", hub_chunks, 0, 0, 0, 2);
                run("hub: 2 LDS per round only", hub_chunks, 0, 0, 0, 3);
            }
        }
        return 0;
    }
    for (int cf = 0; cf < 4; ++cf) {
        pf = cf >= 2 ? 32 : 0;
        if (cf == 2) printf("==== with L2 prefetch (cp.async.bulk.prefetch.L2, 32 hub chunks ahead / next tail chunk)\n");
        fill_words<<<1024, 256>>>(words, (int64_t)hub_chunks * 1024, H, cf & 1);
        fill_cols<<<1024, 256>>>(cols, (int64_t)tail_chunks * 1024, 1 << 21, 1 << 24);
        CK(cudaDeviceSynchronize());
        printf("---- hub words %s; tail columns uniform in [2M, 16M)\n", (cf & 1) ? "conflict-free (lane = bank)" : "random banks");
        run("hub only, LDG.32 words", hub_chunks, 0, 0, 0, 0);
        run("hub only, LDG.128 words", hub_chunks, 0, 0, 0, 1);
        if (cf & 1) continue;
        run("tail only, LSU gathers", 0, tail_chunks, 32, 0, 0);
        run("tail only, LSU gathers, LDG.128 columns", 0, tail_chunks, 32, 2, 0);
        run("tail only, TEX gathers", 0, tail_chunks, 32, 1, 0);
        run("hub + tail LSU, 6 tail warps", hub_chunks, tail_chunks, 6, 0, 0);
        run("hub + tail TEX, 6 tail warps", hub_chunks, tail_chunks, 6, 1, 0);
        run("hub LDG.128 + tail LSU.128, 6 tail warps", hub_chunks, tail_chunks, 6, 2, 1);
        run("hub LDG.128 + tail TEX, 6 tail warps", hub_chunks, tail_chunks, 6, 1, 1);
        run("hub LDG.128 + tail TEX, 10 tail warps", hub_chunks, tail_chunks, 10, 1, 1);
    }
    // L2-resident tail table (columns in a 8 MB window): hit-rate sensitivity of both tail paths
    fill_cols<<<1024, 256>>>(cols, (int64_t)tail_chunks * 1024, 1 << 21, (1 << 21) + (1 << 21));
    CK(cudaDeviceSynchronize());
    printf("---- tail columns uniform in an 8 MB window (L2 hits)\n");
    run("tail only, LSU gathers", 0, tail_chunks, 32, 0, 0);
    run("tail only, TEX gathers", 0, tail_chunks, 32, 1, 0);
    return 0;
}
