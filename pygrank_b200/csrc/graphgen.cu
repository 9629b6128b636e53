// Synthetic edge generators (RMAT, Barabási–Albert-like) — counter-based, so the numpy twins in
// pygrank_b200/synthetic.py produce bit-identical edge lists.  Bench/test utility, no reference
// counterpart (the reference downloads its graphs, pygrank/benchmarks/download.py:62-72).
#include <stdarg.h>

#include "common.cuh"

namespace pgb {

static thread_local char g_error[512] = "";
char *error_buffer() { return g_error; }

int fail(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return 1;
}

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= PGB_MAX_DEVICES) dev = 0;
    return dev;
}

int sm_count() {
    static PerDeviceInt cached;
    int &c = cached.here();
    if (c == 0) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, current_device()) == cudaSuccess && sms > 0)
            c = sms;
        else
            c = 148;
    }
    return c;
}

__global__ void rmat_kernel(int scale, int64_t first_edge, int64_t num_edges, uint64_t seed_hash, uint32_t t1,
                            uint32_t t2, uint32_t t3, int32_t *__restrict__ src, int32_t *__restrict__ dst) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < num_edges;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t e = (uint64_t)(first_edge + i);
        const uint64_t s = mix64(seed_hash ^ (e * 0xD1342543DE82EF95ull));
        uint32_t u_src = 0, u_dst = 0;
        uint64_t h = 0;
        for (int lvl = 0; lvl < scale; ++lvl) {
            uint32_t u;
            if ((lvl & 1) == 0) {
                h = mix64(s + (uint64_t)(lvl >> 1) * 0x9E3779B97F4A7C15ull);
                u = (uint32_t)(h >> 32);
            } else {
                u = (uint32_t)(h & 0xFFFFFFFFull);
            }
            const uint32_t sbit = (u >= t2) ? 1u : 0u;
            const uint32_t dbit = ((u >= t1 && u < t2) || u >= t3) ? 1u : 0u;
            u_src = (u_src << 1) | sbit;
            u_dst = (u_dst << 1) | dbit;
        }
        src[i] = (int32_t)u_src;
        dst[i] = (int32_t)u_dst;
    }
}

__global__ void ba_kernel(int64_t m, int64_t first_slot, int64_t num_slots, uint64_t seed_hash,
                          int32_t *__restrict__ src, int32_t *__restrict__ dst) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < num_slots;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = first_slot + i;
        int64_t cur = k;
        int64_t target;
        for (;;) {
            const int64_t limit = 2 * (cur / m) * m;  // endpoints owned by earlier nodes
            if (limit == 0) {
                target = cur % m;
                break;
            }
            const uint64_t h = mix64(seed_hash ^ ((uint64_t)cur * 0xD1342543DE82EF95ull));
            const int64_t r = (int64_t)(h % (uint64_t)limit);
            const int64_t slot = r >> 1;
            if ((r & 1) == 0) {
                target = slot / m + m;
                break;
            }
            cur = slot;
        }
        src[i] = (int32_t)(k / m + m);
        dst[i] = (int32_t)target;
    }
}

}  // namespace pgb

using namespace pgb;

extern "C" {

int pgb_abi_version(void) { return PGB_ABI_VERSION; }
const char *pgb_last_error(void) { return error_buffer(); }

int pgb_device_sm_count(int device) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    return sms;
}

int pgb_rmat_edges(int scale, int64_t first_edge, int64_t num_edges, uint64_t seed, uint32_t t1, uint32_t t2,
                   uint32_t t3, int32_t *src, int32_t *dst, void *stream) {
    if (scale < 1 || scale > 31) return fail("pgb_rmat_edges: scale %d outside 1..31", scale);
    if (num_edges <= 0) return 0;
    uint64_t z = seed + 0x9E3779B97F4A7C15ull;  // host copy of mix64(seed)
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    rmat_kernel<<<stride_grid(num_edges, 256), 256, 0, as_stream(stream)>>>(scale, first_edge, num_edges, z, t1, t2,
                                                                            t3, src, dst);
    PGB_LAUNCH_OK("rmat_kernel");
    return 0;
}

int pgb_ba_edges(int64_t n, int m, int64_t first_slot, int64_t num_slots, uint64_t seed, int32_t *src,
                 int32_t *dst, void *stream) {
    if (m < 1 || n <= m) return fail("pgb_ba_edges: need n > m >= 1 (n=%lld m=%d)", (long long)n, m);
    if (num_slots <= 0) return 0;
    uint64_t z = seed + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    ba_kernel<<<stride_grid(num_slots, 256), 256, 0, as_stream(stream)>>>(m, first_slot, num_slots, z, src, dst);
    PGB_LAUNCH_OK("ba_kernel");
    return 0;
}

}  // extern "C"
