// Hub-blocked sliced-ELL (hsell) — the per-iteration gather of the hot path for unweighted graphs,
// organised around what the measurements of round 1 showed: a scattered 4-byte gather from
// global memory costs one L1 wavefront (~1 cycle per SM) whether it hits or misses, so a CSR
// row-gather is capped near 270 G gathers/s no matter how little HBM traffic it causes, while a
// shared-memory gather costs ~0.11 cycle (128 B/clk/SM over 32 banks, ~3.5-way conflicts on random
// addresses).  On degree-ranked power-law graphs a few hundred thousand hub columns carry 80-90 % of
// the entries, so:
//
//   * columns are cut into hub blocks of `block_cols` (what fits in the 227 KB of shared memory);
//     a CTA loads one block of the gather vector z with coalesced 16-byte loads and serves every
//     entry of that block from shared memory; entries are stored as 16-bit block-local columns
//     (half the index bytes of CSR);
//   * rows are cut into slices of 32, one lane per row, entries stored [round][lane]: index loads are
//     coalesced, every lane accumulates privately, there is no segmented reduction and no per-row
//     bookkeeping in the inner loop (degree ranking makes neighbouring rows equally long);
//   * what is left (the tail: columns past the hub blocks, and slices too sparse in a block to be
//     worth a unit) is gathered from L2 by "tail units" — on other warps of the same CTA, so the
//     L1-wavefront-bound tail and the shared-memory-bound hub work overlap;
//   * the rounds of all (slice, block) units form one stream per kind, cut into chunks of 32 rounds —
//     the uniform unit of work of a warp (no per-unit latency chain, exact load balance); a piece of
//     a unit (it ends at the unit's last round or at the chunk end) writes one 128-byte row of
//     partial sums; a second light kernel adds the partial rows of each slice and applies the fused
//     filter update + convergence reduction (RowUpdate).
//
// Replaces, per iteration, the same reference sequence as spmv_fused.cu: conv
// (/root/reference/pygrank/core/backend/numpy.py:64-65), the filter formula (algorithms/filters/adhoc.py:34-36,
// 166-169; abstract_filters.py:225-256), the quotient (abstract_filters.py:126-136) and the convergence
// check (algorithms/convergence.py:77-101).
#include <stdlib.h>

#include <mutex>

#include "step_common.cuh"

namespace pgb {

constexpr int HS_THREADS = 1024;
constexpr int HS_WARPS = HS_THREADS / 32;
constexpr int HS_SMEM_LIMIT = 232448;   // 227 KB opt-in dynamic shared memory per CTA on sm_100a
constexpr int UPD_BLOCK = 256;
constexpr int UPD_AHEAD = 5;   // partial rows per slice requested before the first add

static int g_tail_warps = 5;   // measured with TEX-path tail gathers: 4 -> .586, 5 -> .572, 6 -> .563, 8 -> .585 ms (RMAT-24 fp32, box-to-box +-3 %)

__device__ __forceinline__ unsigned ld_stream_u32(const uint32_t *p) { return __ldcs(p); }

// ---------------------------------------------------------------------------------------------
// builders
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int lower_bound_i32(const int32_t *__restrict__ a, int lo, int hi, int key) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// virtual column (what the builders' CSR is indexed by) -> position in the gather vector
__device__ __forceinline__ int32_t hsell_real_col(int32_t v, int H, int K, int N, int64_t seg_len) {
    if (N == 1) return v;
    const int Hs = H / N;
    const int64_t hub_span = (int64_t)K * H;
    if (v < hub_span) {
        const int blk = v / H, local = v - blk * H;
        const int rnk = local / Hs, off = local - rnk * Hs;
        return (int32_t)(rnk * seg_len + (int64_t)blk * Hs + off);
    }
    const int64_t vt = v - hub_span;
    const int64_t tl = seg_len - (int64_t)K * Hs;   // tail entries per segment
    const int64_t rnk = vt / tl, off = vt - rnk * tl;
    return (int32_t)(rnk * seg_len + (int64_t)K * Hs + off);
}

constexpr int HS_MAX_WINDOWS = 16;

// Tail window of a virtual column: the tail of a slice is cut by the position of the gathered entry in the gather
// vector, window w = [w*window_len, (w+1)*window_len).  With the vector of a row-partitioned graph several times the
// 126 MB L2, all CTAs then sweep the tail stream window by window and the L2 keeps ONE window of z (one rank's range)
// instead of thrashing on 32-byte sectors of the whole vector.
__device__ __forceinline__ int hsell_window(int32_t v, int H, int K, int N, int64_t seg_len, int W, int64_t window_len) {
    if (W == 1) return 0;
    const int w = (int)((int64_t)hsell_real_col(v, H, K, N, seg_len) / window_len);
    return w < W ? w : W - 1;
}

__global__ void hsell_count_kernel(int64_t n, int64_t n_slices, const int32_t *__restrict__ indptr,
                                   const int32_t *__restrict__ indices, int H, int K, int min_entries,
                                   float round_cost, int N, int64_t seg_len, int W, int64_t window_len,
                                   int window_min_rounds, int32_t *__restrict__ hub_rounds,
                                   int32_t *__restrict__ tail_rounds) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const unsigned FULL = 0xffffffffu;
    for (int64_t s = warp; s < n_slices; s += nwarps) {
        const int64_t row = s * 32 + lane;
        int b = 0, e = 0;
        if (row < n) {
            b = indptr[row];
            e = indptr[row + 1];
        }
        int pos = b;
        int tail_len[HS_MAX_WINDOWS];
#pragma unroll
        for (int w = 0; w < HS_MAX_WINDOWS; ++w) tail_len[w] = 0;
        auto to_tail = [&](int from, int to) {
            if (W == 1) {
                tail_len[0] += to - from;
                return;
            }
            for (int i = from; i < to; ++i) {
                const int w = hsell_window(indices[i], H, K, N, seg_len, W, window_len);
#pragma unroll
                for (int q = 0; q < HS_MAX_WINDOWS; ++q)
                    if (q == w) ++tail_len[q];
            }
        };
        for (int blk = 0; blk < K; ++blk) {
            const int nxt = lower_bound_i32(indices, pos, e, (blk + 1) * H);
            const int len = nxt - pos;
            const int ent = __reduce_add_sync(FULL, len);
            const int mx = __reduce_max_sync(FULL, len);
            const int rounds = (mx + 1) / 2;
            const bool use = ent > 0 && (float)ent >= (float)min_entries + round_cost * (float)rounds;
            if (lane == 0) hub_rounds[(int64_t)blk * n_slices + s] = use ? rounds : 0;
            if (!use) to_tail(pos, nxt);
            pos = nxt;
        }
        to_tail(pos, e);
        if (W > 1) {
            // a slice whose rows hold only a few tail entries stays ONE unit (in the last window): cut into windows
            // it would pay a padded round per window for a handful of entries; the long tails — hub rows, where
            // nearly all tail entries are — are the ones worth sweeping window by window
            int total = 0;
#pragma unroll
            for (int w = 0; w < HS_MAX_WINDOWS; ++w) total += tail_len[w];
            if (__reduce_max_sync(FULL, total) < window_min_rounds) {
#pragma unroll
                for (int w = 0; w < HS_MAX_WINDOWS; ++w) tail_len[w] = (w == W - 1) ? total : 0;
            }
        }
#pragma unroll
        for (int w = 0; w < HS_MAX_WINDOWS; ++w) {
            if (w < W) {
                const int tmx = __reduce_max_sync(FULL, tail_len[w]);
                if (lane == 0) tail_rounds[(int64_t)w * n_slices + s] = tmx;
            }
        }
    }
}

constexpr int FILL_GROUP = 8;   // hub blocks per work item of the fill kernel

// One warp per (slice, part): part j < NP-1 writes the hub units of blocks [j*FILL_GROUP, (j+1)*FILL_GROUP),
// the last part the slice's tail (entries of the blocks without a unit, then the columns past the hub
// blocks).  Splitting the slices matters for the hub rows: the top slice alone holds ~1 % of all entries
// and used to be the critical path of the whole build.
template <typename VT, bool WEIGHTED>
__global__ void hsell_fill_kernel(int64_t n, int64_t n_slices, const int32_t *__restrict__ indptr,
                                  const int32_t *__restrict__ indices, const VT *__restrict__ values,
                                  VT *__restrict__ hub_vals, VT *__restrict__ tail_vals, int H, int K, int N,
                                  int64_t seg_len,
                                  const int32_t *__restrict__ hub_rounds, const int32_t *__restrict__ tail_rounds,
                                  const int64_t *__restrict__ hub_round_base, const int64_t *__restrict__ hub_part_base,
                                  const int64_t *__restrict__ tail_round_base,
                                  const int64_t *__restrict__ tail_part_base, const int32_t *__restrict__ slice_ptr,
                                  uint32_t *__restrict__ hub_words, int32_t *__restrict__ tail_cols,
                                  int32_t *__restrict__ piece_row, int32_t *__restrict__ scratch, int banks, int W,
                                  int64_t window_len) {
    constexpr int CH = PGB_HSELL_CHUNK;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int NP = (K + FILL_GROUP - 1) / FILL_GROUP + 1;
    for (int64_t item = warp; item < n_slices * NP; item += nwarps) {
        const int64_t s = item / NP;
        const int part = (int)(item - s * NP);
        const int64_t row = s * 32 + lane;
        int b = 0, e = 0;
        if (row < n) {
            b = indptr[row];
            e = indptr[row + 1];
        }
        const int64_t first_part = slice_ptr[s];
        auto pieces_of = [&](int blk) -> int {
            const int R = hub_rounds[(int64_t)blk * n_slices + s];
            if (R <= 0) return 0;
            const int64_t g0 = hub_round_base[(int64_t)blk * n_slices + s];
            return (int)((g0 + R - 1) / CH - g0 / CH) + 1;
        };
        if (part < NP - 1) {
            const int blk_lo = part * FILL_GROUP;
            const int blk_hi = (blk_lo + FILL_GROUP < K) ? blk_lo + FILL_GROUP : K;
            int64_t ord = 0;
            for (int blk = 0; blk < blk_lo; ++blk) ord += pieces_of(blk);
            int pos = lower_bound_i32(indices, b, e, blk_lo * H);
            for (int blk = blk_lo; blk < blk_hi; ++blk) {
                const int nxt = lower_bound_i32(indices, pos, e, (blk + 1) * H);
                const int len = nxt - pos;
                const int R = hub_rounds[(int64_t)blk * n_slices + s];
                if (R > 0) {
                const int64_t g0 = hub_round_base[(int64_t)blk * n_slices + s];
                const int64_t wb = g0 * 32;
                const int base = blk * H;
                // WEIGHTED: the value of the entry in half hh of word (round j, lane) sits at hub_vals[2*word + hh]
                // (a lane reads both halves' values of a round as one 8- or 16-byte load); padding slots keep 0
                if (scratch == nullptr || len < 2) {
                    for (int j = 0; j < R; ++j) {
                        const int i0 = 2 * j, i1 = 2 * j + 1;
                        const uint32_t lo = (i0 < len) ? (uint32_t)(indices[pos + i0] - base) : (uint32_t)H;
                        const uint32_t hi = (i1 < len) ? (uint32_t)(indices[pos + i1] - base) : (uint32_t)H;
                        hub_words[wb + (int64_t)j * 32 + lane] = lo | (hi << 16);
                        if (WEIGHTED) {
                            if (i0 < len) hub_vals[2 * (wb + (int64_t)j * 32 + lane)] = values[pos + i0];
                            if (i1 < len) hub_vals[2 * (wb + (int64_t)j * 32 + lane) + 1] = values[pos + i1];
                        }
                    }
                } else {
                    // Bank-aware slot order.  The 32 lanes of a round read shared memory together; with the
                    // row's entries in column order their banks are random (measured 2.6 wavefronts per
                    // LDS).  Entry position p of lane l is given a column of bank (l + p) mod `banks`
                    // whenever the row has one left, so the lanes of one instruction mostly hit distinct
                    // banks; the columns without a matching position fill the holes.
                    const uint32_t HOLE = 0xffffu;
                    int cnt[32], cur[32];
#pragma unroll
                    for (int q = 0; q < 32; ++q) cnt[q] = 0;
                    for (int i = 0; i < len; ++i) ++cnt[(indices[pos + i] - base) & (banks - 1)];
                    int run = 0;
                    for (int q = 0; q < banks; ++q) {
                        cur[q] = run;
                        run += cnt[q];
                    }
                    for (int i = 0; i < len; ++i) {   // scratch: the row's entries (offsets from pos) bucketed by bank
                        const int c = indices[pos + i] - base;
                        scratch[pos + cur[c & (banks - 1)]++] = i;
                    }
                    // cur[q] is now the END of bucket q; the bucket starts cnt[q] earlier
                    for (int q = 0; q < banks; ++q) cur[q] -= cnt[q];
                    for (int j = 0; j < R; ++j) {
                        uint32_t half[2];
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const int pp = 2 * j + hh;
                            if (pp >= len) {
                                half[hh] = (uint32_t)H;
                            } else {
                                const int tb = (lane + pp) & (banks - 1);
                                if (cnt[tb] > 0) {
                                    const int src = scratch[pos + cur[tb]++];
                                    half[hh] = (uint32_t)(indices[pos + src] - base);
                                    if (WEIGHTED) hub_vals[2 * (wb + (int64_t)j * 32 + lane) + hh] = values[pos + src];
                                    --cnt[tb];
                                } else {
                                    half[hh] = HOLE;
                                }
                            }
                        }
                        hub_words[wb + (int64_t)j * 32 + lane] = half[0] | (half[1] << 16);
                    }
                    int bb = 0;
                    for (int j = 0; j < R; ++j) {
                        uint32_t w = hub_words[wb + (int64_t)j * 32 + lane];
                        bool changed = false;
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            if (((w >> (16 * hh)) & 0xffffu) == HOLE) {
                                while (cnt[bb] == 0) ++bb;
                                const int src = scratch[pos + cur[bb]++];
                                const uint32_t v = (uint32_t)(indices[pos + src] - base);
                                if (WEIGHTED) hub_vals[2 * (wb + (int64_t)j * 32 + lane) + hh] = values[pos + src];
                                --cnt[bb];
                                w = (w & ~(0xffffu << (16 * hh))) | (v << (16 * hh));
                                changed = true;
                            }
                        }
                        if (changed) hub_words[wb + (int64_t)j * 32 + lane] = w;
                    }
                }
                // pieces: the unit is cut at every chunk boundary of its stream
                const int pieces = (int)((g0 + R - 1) / CH - g0 / CH) + 1;
                const int64_t p0 = hub_part_base[(int64_t)blk * n_slices + s];
                for (int p = lane; p < pieces; p += 32) piece_row[p0 + p] = (int32_t)(first_part + ord + p);
                ord += pieces;
                }
                pos = nxt;
            }
            continue;
        }
        // ---- the tail of the slice: one unit per window (blocks without a unit first, then the columns past the
        //      hub blocks; a row's entries keep their column order inside every window) ---------------------
        int t[HS_MAX_WINDOWS];
#pragma unroll
        for (int w = 0; w < HS_MAX_WINDOWS; ++w) t[w] = 0;
        bool collapsed = W > 1;     // the count pass left this slice one tail unit, in the last window
        for (int w = 0; w + 1 < W; ++w) collapsed = collapsed && tail_rounds[(int64_t)w * n_slices + s] == 0;
        auto emit = [&](int from, int to) {
            for (int i = from; i < to; ++i) {
                const int32_t v = indices[i];
                const int w = collapsed ? W - 1 : hsell_window(v, H, K, N, seg_len, W, window_len);
                int tw = 0;
#pragma unroll
                for (int q = 0; q < HS_MAX_WINDOWS; ++q)
                    if (q == w) tw = t[q]++;
                const int64_t base = tail_round_base[(int64_t)w * n_slices + s] * 32;
                tail_cols[base + (int64_t)tw * 32 + lane] = hsell_real_col(v, H, K, N, seg_len);
                if (WEIGHTED) tail_vals[base + (int64_t)tw * 32 + lane] = values[i];
            }
        };
        int pos = b;
        int64_t ord = 0;
        for (int blk = 0; blk < K; ++blk) {
            const int nxt = lower_bound_i32(indices, pos, e, (blk + 1) * H);
            const int np = pieces_of(blk);
            ord += np;
            if (np == 0) emit(pos, nxt);
            pos = nxt;
        }
        emit(pos, e);
        for (int w = 0; w < W; ++w) {
            const int TR = tail_rounds[(int64_t)w * n_slices + s];
            if (TR > 0) {
                const int64_t tg0 = tail_round_base[(int64_t)w * n_slices + s];
                const int pieces = (int)((tg0 + TR - 1) / CH - tg0 / CH) + 1;
                const int64_t p0 = tail_part_base[(int64_t)w * n_slices + s];
                for (int p = lane; p < pieces; p += 32) piece_row[p0 + p] = (int32_t)(first_part + ord + p);
                ord += pieces;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// kernel A: the gather.  One CTA of 32 warps per SM; shared memory holds one hub block of z.
// ---------------------------------------------------------------------------------------------
// graph_dropout inside the gather (K8; semantics of the torch backends' graph_dropout,
// /root/reference/pygrank/core/backend/pytorch.py:34-38: every stored entry is dropped with probability p and the
// survivors are rescaled by 1/(1-p), a fresh mask per call): the Bernoulli draw of an entry is a counter-based hash of
// (seed, step, position of the entry's slot in its stream) — no mask array, no values array, nothing streamed.
struct DropParams {
    uint64_t key;      // mix of the seed and the step: a new mask every iteration
    uint32_t thr;      // drop when the 32 hash bits are below thr = p * 2^32
    float scale;       // 1 / (1 - p), applied when a piece is flushed
};
template <bool DROP>
__device__ __forceinline__ bool kept(const DropParams &D, uint64_t slot) {
    if (!DROP) return true;
    return (uint32_t)(mix64(D.key ^ (slot * 0xD1342543DE82EF95ull)) >> 32) >= D.thr;
}

// ---------------------------------------------------------------------------------------------
// Panel elements.  The same gather kernel runs on a PANEL of seed columns when the element type of z is a 16-byte
// vector — 4 x fp32 or 2 x fp64 per node (z is [n_cols][PB] row-major): one 16-bit hub index (or one 32-bit tail
// index) then fetches PB useful values with one LDS.128 / one 16-byte texel, the index streams, the chunk
// bookkeeping and the hub-block reloads are paid once per panel instead of once per column, and a piece leaves
// as one 16-byte RED per lane (red.global.add.v4.f32 on sm_100a).  Hub blocks hold 8192 nodes (128 KB).
// ---------------------------------------------------------------------------------------------
struct __align__(16) f32x4 {
    float4 v;
    __device__ __forceinline__ f32x4() {}
    __device__ __forceinline__ f32x4(float s) { v = make_float4(s, s, s, s); }
    __device__ __forceinline__ f32x4(float4 q) : v(q) {}
    __device__ __forceinline__ f32x4 &operator+=(const f32x4 &o) {
        v.x += o.v.x; v.y += o.v.y; v.z += o.v.z; v.w += o.v.w;
        return *this;
    }
};
__device__ __forceinline__ f32x4 operator+(f32x4 a, const f32x4 &b) { return a += b; }
__device__ __forceinline__ f32x4 operator*(const f32x4 &a, const f32x4 &b) {
    return f32x4(make_float4(a.v.x * b.v.x, a.v.y * b.v.y, a.v.z * b.v.z, a.v.w * b.v.w));
}
struct __align__(16) f64x2 {
    double2 v;
    __device__ __forceinline__ f64x2() {}
    __device__ __forceinline__ f64x2(double s) { v = make_double2(s, s); }
    __device__ __forceinline__ f64x2(double2 q) : v(q) {}
    __device__ __forceinline__ f64x2 &operator+=(const f64x2 &o) {
        v.x += o.v.x; v.y += o.v.y;
        return *this;
    }
};
__device__ __forceinline__ f64x2 operator+(f64x2 a, const f64x2 &b) { return a += b; }
__device__ __forceinline__ f64x2 operator*(const f64x2 &a, const f64x2 &b) {
    return f64x2(make_double2(a.v.x * b.v.x, a.v.y * b.v.y));
}

// what the gather kernel needs of an element type: a read-only load, a RED, a texel fetch, a texture format
enum { ELEM_F32 = 0, ELEM_F64 = 1, ELEM_F32X4 = 2, ELEM_F64X2 = 3 };
template <typename T> struct Elem;
template <> struct Elem<float> { static constexpr int kind = ELEM_F32; };
template <> struct Elem<double> { static constexpr int kind = ELEM_F64; };
template <> struct Elem<f32x4> { static constexpr int kind = ELEM_F32X4; };
template <> struct Elem<f64x2> { static constexpr int kind = ELEM_F64X2; };

__device__ __forceinline__ float ld_ro(const float *p) { return __ldg(p); }
__device__ __forceinline__ double ld_ro(const double *p) { return __ldg(p); }
__device__ __forceinline__ f32x4 ld_ro(const f32x4 *p) { return f32x4(__ldg(reinterpret_cast<const float4 *>(p))); }
__device__ __forceinline__ f64x2 ld_ro(const f64x2 *p) { return f64x2(__ldg(reinterpret_cast<const double2 *>(p))); }

__device__ __forceinline__ void red_add(float *p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(double *p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(f32x4 *p, const f32x4 &v) { atomicAdd(reinterpret_cast<float4 *>(p), v.v); }
__device__ __forceinline__ void red_add(f64x2 *p, const f64x2 &v) {
    atomicAdd(reinterpret_cast<double *>(p), v.v.x);
    atomicAdd(reinterpret_cast<double *>(p) + 1, v.v.y);
}

// the two values of a weighted hub word (halves 0 and 1) as one streaming load
__device__ __forceinline__ void ld_pair(const float *p, float &a, float &b) {
    const float2 v = __ldcs(reinterpret_cast<const float2 *>(p));
    a = v.x;
    b = v.y;
}
__device__ __forceinline__ void ld_pair(const double *p, double &a, double &b) {
    const double2 v = __ldcs(reinterpret_cast<const double2 *>(p));
    a = v.x;
    b = v.y;
}
template <typename T>
__device__ __forceinline__ void ld_pair(const T *p, T &a, T &b) {   // panel elements: weighted forms are not built
    a = ld_ro(p);
    b = ld_ro(p + 1);
}
__device__ __forceinline__ float ld_val(const float *p) { return __ldcs(p); }
__device__ __forceinline__ double ld_val(const double *p) { return __ldcs(p); }
template <typename T>
__device__ __forceinline__ T ld_val(const T *p) { return ld_ro(p); }

struct GatherParams {
    pgb_hsell h;
    const void *z;
    void *partials;
    const int32_t *stop;   // device state word: run-ahead launches after convergence are no-ops (or NULL)
    uint32_t *tail_queue;  // global counter of the tail stream (zero at launch; the update kernel resets it)
    int tail_warps;
    int tail_batch;        // tail chunks taken per grab of the global queue
    int debug_skip;        // timing experiments only (PGB_HSELL_DEBUG_SKIP): 1 = skip hub chunks, 2 = skip tail chunks
    int bulk;              // hub blocks arrive by cp.async.bulk (TMA unit) + mbarrier instead of LDG -> STS by every thread
    cudaTextureObject_t ztex;   // linear texture over z (TEX kernels): tail gathers go through the TEX pipe
    DropParams drop;            // DROP kernels: in-kernel graph_dropout
};

// A piece's 32 lane sums leave the gather kernel either as one 128-byte partial row (deterministic path: the
// update pass adds the rows of a slice in a fixed order) or — ACCUM — as one coalesced 128-byte RED.ADD into the
// slice's row of an accumulator y kept in L2.  Measured on B200 (scripts/red_rate.cu): a coalesced fp32 RED costs
// what the store costs (2.1 M rows in 0.059 ms either way, hot rows included), fp64 1.5x; and the update pass then
// streams y once instead of chasing 4 partial rows per slice through upd_rows, with no reduce kernel in between.
template <bool ACCUM, typename T>
__device__ __forceinline__ void flush_piece(T *__restrict__ dst, int row, int lane, T v) {
    if (ACCUM)
        red_add(dst + (int64_t)row * 32 + lane, v);
    else
        dst[(int64_t)row * 32 + lane] = v;
}

constexpr int CH = PGB_HSELL_CHUNK;   // rounds per chunk
constexpr int BATCH = 8;              // rounds loaded ahead per lane (scalar elements)
static_assert(CH == 32 && CH % BATCH == 0, "the end mask of a chunk is one 32-bit word");
// 16-byte elements keep 4 rounds in flight: the same bytes per lane, and the gathered values still fit the 64
// registers a 1024-thread CTA leaves each thread
template <typename T> struct BatchOf { static constexpr int value = sizeof(T) > 8 ? 4 : BATCH; };

// One chunk of the hub stream: 32 rounds, two shared-memory gathers per lane and round.  WEIGHTED: every entry is
// multiplied by its edge value (one 8- / 16-byte load per lane and round next to the index word; 4 rounds in flight
// instead of 8 to stay inside the 64 registers of a 1024-thread CTA).
template <typename T, bool ACCUM, bool DROP, bool WEIGHTED>
__device__ __forceinline__ void hub_chunk(const uint32_t *__restrict__ words, const T *__restrict__ vals, int64_t chunk,
                                          uint32_t p_first, uint32_t endmask, const int32_t *__restrict__ piece_row,
                                          const T *s_z, T *__restrict__ partials, int lane, const DropParams &D) {
    constexpr int HB = WEIGHTED ? 4 : BATCH;
    const T dscale = DROP ? (T)D.scale : (T)1;
    const uint64_t slot0 = ((uint64_t)chunk * (CH * 32) + (uint64_t)lane) * 2;   // + round * 64 + half
    const uint32_t *d = words + chunk * (CH * 32) + lane;
    const T *dv = WEIGHTED ? vals + 2 * (chunk * (CH * 32) + lane) : nullptr;      // + round * 64
    endmask |= 0x80000000u;   // the chunk end closes the last piece
    // a chunk has at most 32 pieces: lane j fetches the partial row of piece j
    const int my_row = (lane < __popc(endmask)) ? __ldg(piece_row + p_first + lane) : 0;
    uint32_t w[HB], nx[HB];
    T v0[WEIGHTED ? HB : 1], v1[WEIGHTED ? HB : 1], n0[WEIGHTED ? HB : 1], n1[WEIGHTED ? HB : 1];
#pragma unroll
    for (int u = 0; u < HB; ++u) {
        w[u] = ld_stream_u32(d + u * 32);
        if (WEIGHTED) ld_pair(dv + u * 64, v0[u], v1[u]);
    }
    T a0 = (T)0, a1 = (T)0;
    int p = 0;
#pragma unroll
    for (int bt = 0; bt < CH / HB; ++bt) {
        if (bt + 1 < CH / HB) {
#pragma unroll
            for (int u = 0; u < HB; ++u) {
                nx[u] = ld_stream_u32(d + ((bt + 1) * HB + u) * 32);
                if (WEIGHTED) ld_pair(dv + ((bt + 1) * HB + u) * 64, n0[u], n1[u]);
            }
        }
        const uint32_t m8 = (endmask >> (bt * HB)) & ((1u << HB) - 1u);
        if (m8 == 0u) {
#pragma unroll
            for (int u = 0; u < HB; ++u) {
                const uint64_t sl = slot0 + (uint64_t)(bt * HB + u) * 64;
                T x0 = s_z[w[u] & 0xffffu], x1 = s_z[w[u] >> 16];
                if (WEIGHTED) {
                    x0 = x0 * v0[u];
                    x1 = x1 * v1[u];
                }
                a0 += kept<DROP>(D, sl) ? x0 : (T)0;
                a1 += kept<DROP>(D, sl + 1) ? x1 : (T)0;
            }
        } else {
#pragma unroll
            for (int u = 0; u < HB; ++u) {
                const uint64_t sl = slot0 + (uint64_t)(bt * HB + u) * 64;
                T x0 = s_z[w[u] & 0xffffu], x1 = s_z[w[u] >> 16];
                if (WEIGHTED) {
                    x0 = x0 * v0[u];
                    x1 = x1 * v1[u];
                }
                a0 += kept<DROP>(D, sl) ? x0 : (T)0;
                a1 += kept<DROP>(D, sl + 1) ? x1 : (T)0;
                if ((m8 >> u) & 1u) {
                    flush_piece<ACCUM>(partials, __shfl_sync(0xffffffffu, my_row, p), lane, (a0 + a1) * dscale);
                    ++p;
                    a0 = a1 = (T)0;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            w[u] = nx[u];
            if (WEIGHTED) {
                v0[u] = n0[u];
                v1[u] = n1[u];
            }
        }
    }
}

// z[c] through the texture unit.  Measured on B200 (scripts/gather_paths.cu): a scattered 4-byte gather costs
// about one L1TEX cycle per lane on either input pipe, but the TEX pipe and the LSU pipe (which also carries
// every shared-memory gather and the index streams) run side by side — so the tail's L2 gathers stop competing
// with the hub path for LSU wavefronts.
__device__ __forceinline__ float tex_fetch(cudaTextureObject_t t, int c, float) { return tex1Dfetch<float>(t, c); }
__device__ __forceinline__ double tex_fetch(cudaTextureObject_t t, int c, double) {
    const int2 v = tex1Dfetch<int2>(t, c);
    return __hiloint2double(v.y, v.x);
}
__device__ __forceinline__ f32x4 tex_fetch(cudaTextureObject_t t, int c, const f32x4 &) {
    return f32x4(tex1Dfetch<float4>(t, c));
}
__device__ __forceinline__ f64x2 tex_fetch(cudaTextureObject_t t, int c, const f64x2 &) {
    const int4 v = tex1Dfetch<int4>(t, c);
    return f64x2(make_double2(__hiloint2double(v.y, v.x), __hiloint2double(v.w, v.z)));
}

// One chunk of the tail stream: 32 rounds, one L2 gather per lane and round (padding lanes are off).
template <typename T, bool TEX, bool ACCUM, bool DROP, bool WEIGHTED>
__device__ __forceinline__ void tail_chunk(const int32_t *__restrict__ cols, const T *__restrict__ vals, int64_t chunk,
                                           uint32_t p_first, uint32_t endmask, const int32_t *__restrict__ piece_row,
                                           const T *__restrict__ z, cudaTextureObject_t ztex,
                                           T *__restrict__ partials, int lane, const DropParams &D) {
    constexpr int B = (WEIGHTED && sizeof(T) == 8) ? 4 : BatchOf<T>::value;
    const T dscale = DROP ? (T)D.scale : (T)1;
    const uint64_t tslot0 = (1ull << 62) + (uint64_t)chunk * (CH * 32) + (uint64_t)lane;   // + round * 32
    const int32_t *d = cols + chunk * (CH * 32) + lane;
    endmask |= 0x80000000u;
    const int my_row = (lane < __popc(endmask)) ? __ldg(piece_row + p_first + lane) : 0;
    int32_t c[B], nx[B];
#pragma unroll
    for (int u = 0; u < B; ++u) c[u] = ld_stream(d + u * 32);
    T a0 = (T)0, a1 = (T)0;
    int p = 0;
#pragma unroll
    for (int bt = 0; bt < CH / B; ++bt) {
        T x[B];
#pragma unroll
        for (int u = 0; u < B; ++u) {
            const bool on = c[u] >= 0 && kept<DROP>(D, tslot0 + (uint64_t)(bt * B + u) * 32);
            if (TEX)
                x[u] = on ? tex_fetch(ztex, c[u], T()) : (T)0;
            else
                x[u] = on ? ld_ro(z + c[u]) : (T)0;
            if (WEIGHTED) x[u] = x[u] * ld_val(vals + chunk * (CH * 32) + lane + (bt * B + u) * 32);   // padding: 0 * 0
        }
        if (bt + 1 < CH / B) {
#pragma unroll
            for (int u = 0; u < B; ++u) nx[u] = ld_stream(d + ((bt + 1) * B + u) * 32);
        }
        const uint32_t m8 = (endmask >> (bt * B)) & ((1u << B) - 1u);
        if (m8 == 0u) {
#pragma unroll
            for (int u = 0; u < B; u += 2) {
                a0 += x[u];
                a1 += x[u + 1];
            }
        } else {
#pragma unroll
            for (int u = 0; u < B; ++u) {
                a0 += x[u];
                if ((m8 >> u) & 1u) {
                    flush_piece<ACCUM>(partials, __shfl_sync(0xffffffffu, my_row, p), lane, (a0 + a1) * dscale);
                    ++p;
                    a0 = a1 = (T)0;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < B; ++u) c[u] = nx[u];
    }
}

// Bulk copy of a hub block: one thread hands the 128 KB to the TMA unit (cp.async.bulk, SASS UBLKCP) and arms an
// mbarrier with the byte count; warps that gather from the block wait on the barrier, warps with tail work do not —
// the block load no longer occupies 1024 threads' LDG -> STS slots and overlaps the tail chunks.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_global, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_global), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

template <typename T, bool TEX, bool ACCUM, bool DROP, bool WEIGHTED = false>
__global__ void __launch_bounds__(HS_THREADS, 1) hsell_gather_kernel(const GatherParams G) {
    extern __shared__ __align__(16) unsigned char hs_smem[];
    T *s_z = reinterpret_cast<T *>(hs_smem);   // [block_cols + 1]; the last entry is the padding target (0)
    __shared__ int s_hub_next;
    __shared__ __align__(8) uint64_t s_bar;    // completion of the bulk copy of the current hub block

    if (G.stop && *G.stop != PGB_RUNNING) return;
    const pgb_hsell &h = G.h;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    const T *__restrict__ z = (const T *)G.z;
    T *__restrict__ partials = (T *)G.partials;   // partial rows, or (ACCUM) the accumulator y [n_slices + 1][32]
    const int32_t *__restrict__ piece_dst = ACCUM ? h.piece_slice : h.piece_row;
    const int cta = blockIdx.x;
    const int hub_lo = h.cta_hub_begin[cta], hub_hi = h.cta_hub_begin[cta + 1];
    // hub chunks need this CTA's shared-memory block, so they are dealt statically (block-major ranges);
    // tail chunks need nothing, so every warp of the grid takes them from one global queue: CTAs whose
    // hub share is cheaper absorb more of the tail and all CTAs end together
    const int tail_hi = h.n_tail_chunks;
    const bool tail_pref = warp < G.tail_warps;
    const int H = h.block_cols, N = h.n_segments;
    const int Hs = H / N;
    if (tid == 0) {
        s_z[H] = (T)0;
        if (G.bulk) mbar_init(&s_bar, 1);
    }
    uint32_t bar_parity = 0;   // phase of s_bar the next bulk-loaded block completes (same in every thread)

    const int TB = G.tail_batch;
    auto run_tail = [&](int u0) {   // a grab of the global queue: TB consecutive tail chunks
        if (G.debug_skip & 2) return;
        const int u1 = (u0 + TB < tail_hi) ? u0 + TB : tail_hi;
        for (int u = u0; u < u1; ++u) {
            const uint2 d = __ldg(reinterpret_cast<const uint2 *>(h.tail_chunks) + u);
            tail_chunk<T, TEX, ACCUM, DROP, WEIGHTED>(h.tail_cols, (const T *)h.tail_vals, u, d.x, d.y, piece_dst, z, G.ztex,
                                                      partials, lane, G.drop);
        }
    };

    int cur = hub_lo;
    int blk = 0;
    while (cur < hub_hi) {
        // ---- one segment: the chunks of one hub block inside this CTA's range ------------------------
        while (h.block_chunk_begin[blk + 1] <= cur) ++blk;
        const int blk_end = h.block_chunk_begin[blk + 1];
        const int seg_end = blk_end < hub_hi ? blk_end : hub_hi;
        __syncthreads();   // every warp is done with the previous block
        // bulk path: every segment part 16-byte aligned on both sides and a multiple of 16 bytes long
        constexpr int V = 16 / sizeof(T);
        bool use_bulk = G.bulk != 0;
        uint32_t bulk_bytes = 0;
        for (int sgm = 0; sgm < N && use_bulk; ++sgm) {
            const int64_t first = (int64_t)sgm * h.seg_len + (int64_t)blk * Hs;
            int64_t avail = h.seg_len - (int64_t)blk * Hs;
            const int cnt = (int)(avail < Hs ? (avail < 0 ? 0 : avail) : Hs);
            if ((first % V) || ((sgm * Hs) % V) || (cnt % V) || (((uintptr_t)z) & 15u)) use_bulk = false;
            bulk_bytes += (uint32_t)cnt * (uint32_t)sizeof(T);
        }
        if (bulk_bytes == 0) use_bulk = false;
        if (use_bulk && tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of s_z before the async writes
            mbar_expect_tx(&s_bar, bulk_bytes);
        }
        for (int sgm = 0; sgm < N; ++sgm) {
            const int64_t first = (int64_t)sgm * h.seg_len + (int64_t)blk * Hs;
            int64_t avail = h.seg_len - (int64_t)blk * Hs;
            const int cnt = (int)(avail < Hs ? (avail < 0 ? 0 : avail) : Hs);
            T *dst = s_z + sgm * Hs;
            const T *src = z + first;
            if (use_bulk) {
                if (tid == 0 && cnt > 0) bulk_g2s(dst, src, (uint32_t)cnt * (uint32_t)sizeof(T), &s_bar);
            } else if (((first % V) == 0) && (((sgm * Hs) % V) == 0)) {   // vector part when both sides are 16-byte aligned
                const int nv = cnt / V;
                const float4 *s4 = reinterpret_cast<const float4 *>(src);
                float4 *d4 = reinterpret_cast<float4 *>(dst);
                for (int i = tid; i < nv; i += HS_THREADS) d4[i] = __ldg(s4 + i);
                for (int i = nv * V + tid; i < cnt; i += HS_THREADS) dst[i] = ld_ro(src + i);
            } else {
                for (int i = tid; i < cnt; i += HS_THREADS) dst[i] = ld_ro(src + i);
            }
            for (int i = cnt + tid; i < Hs; i += HS_THREADS) dst[i] = (T)0;
        }
        bool block_ready = !use_bulk;   // bulk: the first hub chunk of every warp waits for the copy
        const uint32_t wait_parity = bar_parity;
        if (use_bulk) bar_parity ^= 1u;
        if (tid == 0) s_hub_next = cur;
        __syncthreads();
        // the id of the NEXT chunk is requested (lane 0) before the current one is processed, so the atomic's
        // round trip — ~1 us on the contended global tail counter — hides behind the chunk's own work
        auto grab = [&](int &kind, int &u) {   // 0: nothing left in this segment, 1: hub chunk, 2: tail chunk
            kind = 0;
            u = 0;
            if (lane == 0) {
                if (tail_pref && *(volatile int *)&s_hub_next < seg_end) {
                    u = (int)atomicAdd(G.tail_queue, (unsigned)TB);
                    if (u < tail_hi) kind = 2;
                }
                if (kind == 0) {
                    u = atomicAdd(&s_hub_next, 1);
                    if (u < seg_end) kind = 1;
                }
            }
        };
        int kind, u;
        grab(kind, u);
        kind = __shfl_sync(FULL, kind, 0);
        u = __shfl_sync(FULL, u, 0);
        while (kind != 0) {
            int nkind, nu;
            grab(nkind, nu);
            if (kind == 1) {
                if (!block_ready) {
                    mbar_wait(&s_bar, wait_parity);
                    block_ready = true;
                }
                if (!(G.debug_skip & 1)) {
                    const uint2 d = __ldg(reinterpret_cast<const uint2 *>(h.hub_chunks) + u);
                    hub_chunk<T, ACCUM, DROP, WEIGHTED>(h.hub_words, (const T *)h.hub_vals, u, d.x, d.y, piece_dst, s_z,
                                                        partials, lane, G.drop);
                }
            } else {
                run_tail(u);
            }
            kind = __shfl_sync(FULL, nkind, 0);
            u = __shfl_sync(FULL, nu, 0);
        }
        cur = seg_end;
    }
    __syncthreads();
    // ---- drain the tail queue ---------------------------------------------------------------------
    {
        int u = 0;
        if (lane == 0) u = (int)atomicAdd(G.tail_queue, (unsigned)TB);
        u = __shfl_sync(FULL, u, 0);
        while (u < tail_hi) {
            int nu = 0;
            if (lane == 0) nu = (int)atomicAdd(G.tail_queue, (unsigned)TB);
            run_tail(u);
            u = __shfl_sync(FULL, nu, 0);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// kernel B: add the partial rows of each slice, fused filter update, convergence reduction.
// Slices with many parts (hub rows: their units span many chunks) are reduced by a whole CTA first,
// in a fixed order (deterministic); the others by one warp each, 32 slices per warp pass.
// ---------------------------------------------------------------------------------------------
constexpr int UPD_WARPS = UPD_BLOCK / 32;

// Kernel B1 (only when some slice has more than heavy_parts pieces): every group of <= 32 consecutive
// partial rows of such a slice is added by one warp into a second-level row, in order (deterministic);
// upd_rows of the slice names the second-level rows.
template <typename T>
__global__ void __launch_bounds__(256) hsell_reduce_kernel(const pgb_hsell h, T *__restrict__ partials,
                                                           const int32_t *stop) {
    if (stop && *stop != PGB_RUNNING) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t it = warp; it < h.n_reduce; it += nwarps) {
        const int32_t start = h.reduce_items[it * 3], cnt = h.reduce_items[it * 3 + 1], out = h.reduce_items[it * 3 + 2];
        const T *src = partials + (int64_t)start * 32 + lane;
        T x[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) x[u] = (u < cnt) ? ld_stream(src + u * 32) : (T)0;
        T acc = (T)0;
#pragma unroll
        for (int u = 0; u < 32; ++u) acc += x[u];
        partials[(int64_t)out * 32 + lane] = acc;
    }
}

// Kernel B: add the partial rows of each slice (one contiguous run, see upd_rows), fused filter
// update, convergence reduction.  Slices that still have more than heavy_parts rows are added by a
// whole CTA first, in a fixed order; the others by one warp, UPD_GROUP slices at a time with all
// their loads in flight together.
template <typename T, int MODE, bool SYMDEG, int UPD_GROUP>
__global__ void __launch_bounds__(UPD_BLOCK, (sizeof(T) == 8 ? 2 : 3) * (UPD_GROUP == 2 ? 2 : 1)) hsell_update_kernel(const StepParams P, const pgb_hsell h,
                                                                                          const T *__restrict__ partials) {
    __shared__ double s_red[32];
    __shared__ T s_acc[UPD_WARPS][32];
    if (MODE != MODE_CONV) {
        if (P.si[PGB_SI_STOP] != PGB_RUNNING) return;
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    const int2 *__restrict__ upd_rows = reinterpret_cast<const int2 *>(h.upd_rows);
    const int heavy_parts = h.heavy_parts;
    RowUpdate<T, MODE, SYMDEG> update(P);
    using Loaded = typename RowUpdate<T, MODE, SYMDEG>::Loaded;

    // ---- heavy slices: one CTA each --------------------------------------------------------------
    for (int hi = blockIdx.x; hi < h.n_heavy; hi += gridDim.x) {
        const int64_t s = h.heavy_slices[hi];
        const int2 rc = upd_rows[s];
        T acc = (T)0;
        for (int j0 = wib * 8; j0 < rc.y; j0 += UPD_WARPS * 8) {
            T x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                x[u] = (j0 + u < rc.y) ? ld_stream(partials + (int64_t)(rc.x + j0 + u) * 32 + lane) : (T)0;
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += x[u];
        }
        s_acc[wib][lane] = acc;
        __syncthreads();
        if (wib == 0) {
            T tot = (T)0;
#pragma unroll
            for (int w = 0; w < UPD_WARPS; ++w) tot += s_acc[w][lane];
            const int64_t row = s * 32 + lane;
            if (row < P.n) {
                int deg = 0;
                if (SYMDEG) deg = P.indptr[row + 1] - P.indptr[row];
                update(row, tot, deg);
            }
        }
        __syncthreads();
    }

    // ---- the other slices: a warp takes UPD_GROUP consecutive slices; the first UPD_AHEAD rows of each
    //      and everything the update reads are requested together (one exposed round trip) -------------
    const int64_t warp = blockIdx.x * (int64_t)UPD_WARPS + wib;
    const int64_t stride = (int64_t)gridDim.x * UPD_WARPS * UPD_GROUP;
    const int64_t n_slices = h.n_slices;
    auto load_rc = [&](int64_t s) -> int2 {
        return (lane < UPD_GROUP && s + lane < n_slices) ? __ldg(upd_rows + s + lane) : make_int2(0, -1);
    };
    update.batch_sums = true;
    int64_t s0 = warp * UPD_GROUP;
    int2 rc_cur = load_rc(s0);
    for (; s0 < n_slices; s0 += stride) {
        const int2 rc_nxt = load_rc(s0 + stride);   // the next group's runs: one iteration ahead
        int cnt[UPD_GROUP];
        uint32_t prow[UPD_GROUP];   // element offset of the slice's first partial row (+ lane): 32-bit on purpose
        bool live[UPD_GROUP];
        T x[UPD_GROUP][UPD_AHEAD];
        Loaded L[UPD_GROUP];
        int ip0[UPD_GROUP], ip1[UPD_GROUP];
#pragma unroll
        for (int k = 0; k < UPD_GROUP; ++k) {
            cnt[k] = __shfl_sync(FULL, rc_cur.y, k);
            live[k] = cnt[k] >= 0 && cnt[k] <= heavy_parts;   // else: past the end / added by a CTA above
            if (!live[k]) cnt[k] = 0;
            prow[k] = (uint32_t)__shfl_sync(FULL, rc_cur.x, k) * 32u + (uint32_t)lane;
#pragma unroll
            for (int u = 0; u < UPD_AHEAD; ++u) x[k][u] = (u < cnt[k]) ? ld_stream(partials + (prow[k] + u * 32u)) : (T)0;
        }
        const int64_t row0 = s0 * 32 + lane;
#pragma unroll
        for (int k = 0; k < UPD_GROUP; ++k) {
            const int64_t row = row0 + k * 32;
            live[k] = live[k] && row < P.n;
            ip0[k] = ip1[k] = 0;
            if (live[k]) {
                if (SYMDEG && MODE != MODE_CONV) {
                    ip0[k] = P.indptr[row];
                    ip1[k] = P.indptr[row + 1];
                }
                L[k] = update.load(row, 0);
            }
        }
#pragma unroll
        for (int k = 0; k < UPD_GROUP; ++k) {
            T acc = x[k][0];
#pragma unroll
            for (int u = 1; u < UPD_AHEAD; ++u) acc += x[k][u];
            for (int t = UPD_AHEAD; t < cnt[k]; t += 4) {   // slices with units in many blocks
                T y[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) y[u] = (t + u < cnt[k]) ? ld_stream(partials + (prow[k] + (uint32_t)(t + u) * 32u)) : (T)0;
                acc += (y[0] + y[1]) + (y[2] + y[3]);
            }
            if (live[k]) {
                update.set_degree(L[k], ip1[k] - ip0[k]);
                update.apply(row0 + k * 32, acc, L[k]);
            }
        }
        update.flush_batch();
        rc_cur = rc_nxt;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0 && P.span_cnt) P.span_cnt[0] = 0u;   // tail queue of the next gather
    if (MODE != MODE_CONV) step_epilogue(P, update, s_red);
}

// Kernel B, accumulate mode: the gathered sums sit complete in y (one value per row), so the update pass is a pure
// stream — y (read, then zeroed for the next step), the row pointers, z, q, c in, z' out — four rows per thread with
// all their loads requested together, then the fused filter update and the convergence reduction as above.
constexpr int UPA_BLOCK = 256;
constexpr int UPA_ROWS = 4;
template <typename T, int MODE, bool SYMDEG>
__global__ void __launch_bounds__(UPA_BLOCK) hsell_update_accum_kernel(const StepParams P, T *__restrict__ y) {
    __shared__ double s_red[32];
    if (MODE != MODE_CONV) {
        if (P.si[PGB_SI_STOP] != PGB_RUNNING) return;
    }
    RowUpdate<T, MODE, SYMDEG> update(P);
    using Loaded = typename RowUpdate<T, MODE, SYMDEG>::Loaded;
    update.batch_sums = true;
    const int64_t n = P.n;
    const int64_t stride = (int64_t)gridDim.x * UPA_BLOCK;
    for (int64_t base = blockIdx.x * (int64_t)UPA_BLOCK + threadIdx.x; base < n; base += stride * UPA_ROWS) {
        T acc[UPA_ROWS];
        Loaded L[UPA_ROWS];
        int ip0[UPA_ROWS], ip1[UPA_ROWS];
#pragma unroll
        for (int k = 0; k < UPA_ROWS; ++k) {
            const int64_t row = base + k * stride;
            acc[k] = (T)0;
            ip0[k] = ip1[k] = 0;
            if (row < n) {
                acc[k] = ld_stream(y + row);
                if (SYMDEG && MODE != MODE_CONV) {
                    ip0[k] = P.indptr[row];
                    ip1[k] = P.indptr[row + 1];
                }
                L[k] = update.load(row, 0);
            }
        }
#pragma unroll
        for (int k = 0; k < UPA_ROWS; ++k) {
            const int64_t row = base + k * stride;
            if (row < n) {
                y[row] = (T)0;
                update.set_degree(L[k], ip1[k] - ip0[k]);
                update.apply(row, acc[k], L[k]);
            }
        }
        update.flush_batch();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0 && P.span_cnt) P.span_cnt[0] = 0u;   // tail queue of the next gather
    if (MODE != MODE_CONV) step_epilogue(P, update, s_red);
}

// ---------------------------------------------------------------------------------------------
// Kernel B of the PANEL path (K3 on the hub-blocked form): PB seed columns advance together, every column with its
// own alpha, normaliser, error sum and stop decision (finalize_column) — NodeRanking.propagate
// (/root/reference/pygrank/core/signals.py:225-226) runs one solve per column; a frozen column keeps z' = z from the
// iteration at which the reference's per-column loop would have stopped.  One thread per row and pass: the row's PB
// gathered sums (y, then zeroed), z, q move as 16-byte vectors, the per-row factors once for all columns.
// ---------------------------------------------------------------------------------------------
struct PanelStep {
    int64_t n;
    const int32_t *indptr;   // SYMDEG: degree-derived w / sq
    const void *zin, *q;     // [n][PB]
    void *zout;              // [n][PB]
    const void *w, *sq, *c;  // per row
    void *y;                 // [(n_slices + 1) * 32][PB]
    double *sf;              // [PB][PGB_STATE_F64_LEN]
    int32_t *si;             // [PB][PGB_STATE_I32_LEN], then the shared ticket, the panel stop word, the step counter
    double *err_hist;        // [PB][hist_stride]
    int32_t hist_stride;
    uint32_t *tail_queue;
    void *ranks;             // MODE_POLY: accumulated results [n][PB]
    const double *coef;      // MODE_POLY: coefficient of step k at coef[k]
};

template <typename S, int PB> struct alignas(16) Pack { S v[PB]; };
__device__ __forceinline__ Pack<float, 4> ld_pack_rw(const float *p) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    Pack<float, 4> r;
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
}
__device__ __forceinline__ Pack<double, 2> ld_pack_rw(const double *p) {
    const double2 t = *reinterpret_cast<const double2 *>(p);
    Pack<double, 2> r;
    r.v[0] = t.x; r.v[1] = t.y;
    return r;
}
__device__ __forceinline__ Pack<float, 4> ld_pack(const float *p) {
    const float4 t = __ldcs(reinterpret_cast<const float4 *>(p));
    Pack<float, 4> r;
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
}
__device__ __forceinline__ Pack<double, 2> ld_pack(const double *p) {
    const double2 t = __ldcs(reinterpret_cast<const double2 *>(p));
    Pack<double, 2> r;
    r.v[0] = t.x; r.v[1] = t.y;
    return r;
}
__device__ __forceinline__ void st_pack(float *p, const Pack<float, 4> &r) {
    *reinterpret_cast<float4 *>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
}
__device__ __forceinline__ void st_pack(double *p, const Pack<double, 2> &r) {
    *reinterpret_cast<double2 *>(p) = make_double2(r.v[0], r.v[1]);
}

constexpr int UPP_BLOCK = 256;
constexpr int UPP_ROWS = 2;
template <typename S, int PB, bool SYMDEG, int MODE>
__global__ void __launch_bounds__(UPP_BLOCK) hsell_update_panel_kernel(const PanelStep P) {
    __shared__ double s_red[2][UPP_BLOCK / 32][PB];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bool active[PB];
    S invS[PB], alpha[PB];
    bool any = false;
    int err_mode = PGB_ERR_MABS;   // one measure per panel: taken from a running column (empty slots hold no settings)
#pragma unroll
    for (int c = 0; c < PB; ++c) {
        active[c] = P.si[c * PGB_STATE_I32_LEN + PGB_SI_STOP] == PGB_RUNNING;
        invS[c] = (S)P.sf[c * PGB_STATE_F64_LEN + PGB_SF_INVS];
        alpha[c] = (S)P.sf[c * PGB_STATE_F64_LEN + PGB_SF_AMUL];
        if (active[c]) err_mode = P.si[c * PGB_STATE_I32_LEN + PGB_SI_ERR_MODE];
        any = any || active[c];
    }
    if (!any) return;   // run-ahead launch after every column has stopped
    S coef[PB];         // MODE_POLY: every column is at its own step of the coefficient table
#pragma unroll
    for (int c = 0; c < PB; ++c)
        coef[c] = (MODE == MODE_POLY && active[c]) ? (S)P.coef[P.si[c * PGB_STATE_I32_LEN + PGB_SI_STEPS] + 1] : (S)0;
    S *__restrict__ ranks = (S *)P.ranks;
    const bool is_max = err_mode == PGB_ERR_MAX;
    double err[PB], tsum[PB];
#pragma unroll
    for (int c = 0; c < PB; ++c) err[c] = tsum[c] = 0.0;
    const S *__restrict__ zin = (const S *)P.zin;
    const S *__restrict__ qv = (const S *)P.q;
    S *__restrict__ zout = (S *)P.zout;
    S *__restrict__ y = (S *)P.y;
    const int64_t n = P.n;
    const int64_t stride = (int64_t)gridDim.x * UPP_BLOCK;
    for (int64_t base = blockIdx.x * (int64_t)UPP_BLOCK + tid; base < n; base += stride * UPP_ROWS) {
        Pack<S, PB> acc[UPP_ROWS], zi[UPP_ROWS], qi[UPP_ROWS];
        S wi[UPP_ROWS], sqi[UPP_ROWS], ci[UPP_ROWS];
        int ip0[UPP_ROWS], ip1[UPP_ROWS];
#pragma unroll
        for (int k = 0; k < UPP_ROWS; ++k) {
            const int64_t row = base + k * stride;
            ip0[k] = ip1[k] = 0;
            wi[k] = sqi[k] = ci[k] = (S)0;
            if (row < n) {
                acc[k] = ld_pack(y + row * PB);
                zi[k] = ld_pack(zin + row * PB);
                if (MODE == MODE_AFFINE) {
                    qi[k] = ld_pack(qv + row * PB);
                    ci[k] = ld_stream((const S *)P.c + row);
                } else {
                    qi[k] = ld_pack_rw(ranks + row * PB);   // the accumulated result (read-modify-write)
                }
                if (SYMDEG) {
                    ip0[k] = P.indptr[row];
                    ip1[k] = P.indptr[row + 1];
                } else {
                    wi[k] = ld_stream((const S *)P.w + row);
                    sqi[k] = ld_stream((const S *)P.sq + row);
                }
            }
        }
        S berr[PB], bt[PB];
#pragma unroll
        for (int c = 0; c < PB; ++c) berr[c] = bt[c] = (S)0;
#pragma unroll
        for (int k = 0; k < UPP_ROWS; ++k) {
            const int64_t row = base + k * stride;
            if (row < n) {
                if (SYMDEG) {
                    const int deg = ip1[k] - ip0[k];
                    wi[k] = deg > 0 ? RowMath<S>::inv((S)deg) : (S)0;
                    sqi[k] = deg > 0 ? RowMath<S>::root((S)deg) : (S)1;
                }
                Pack<S, PB> zn, zero, rk;
                bool rk_dirty = false;
#pragma unroll
                for (int c = 0; c < PB; ++c) {
                    zero.v[c] = (S)0;
                    zn.v[c] = zi[k].v[c];
                    rk.v[c] = qi[k].v[c];
                    if (!active[c]) continue;
                    if (MODE == MODE_AFFINE) {
                        const S znew = (alpha[c] * wi[k] * acc[k].v[c] + qi[k].v[c]) * invS[c];
                        zn.v[c] = znew;
                        S d = znew - zi[k].v[c];
                        d = sqi[k] * (d < (S)0 ? -d : d);
                        if (is_max)
                            berr[c] = berr[c] > d ? berr[c] : d;
                        else
                            berr[c] += (err_mode == PGB_ERR_MSQ) ? d * d : d;
                        bt[c] += znew * ci[k];
                    } else {   // ClosedFormGraphFilter._step (abstract_filters.py:225-228, 248-256), as RowUpdate MODE_POLY
                        const S pw = sqi[k] * zi[k].v[c];
                        if (coef[c] != (S)0) {
                            const S prev = qi[k].v[c];
                            const S cur = prev + pw * coef[c];
                            rk.v[c] = cur;
                            rk_dirty = true;
                            S d = (sizeof(S) == 8) ? prev - cur : coef[c] * pw;
                            d = d < (S)0 ? -d : d;
                            if (is_max)
                                berr[c] = berr[c] > d ? berr[c] : d;
                            else
                                berr[c] += (err_mode == PGB_ERR_MSQ) ? d * d : d;
                        }
                        zn.v[c] = wi[k] * acc[k].v[c];
                    }
                }
                st_pack(y + row * PB, zero);
                st_pack(zout + row * PB, zn);
                if (MODE == MODE_POLY && rk_dirty) st_pack(ranks + row * PB, rk);
            }
        }
#pragma unroll
        for (int c = 0; c < PB; ++c) {
            if (is_max)
                err[c] = err[c] > (double)berr[c] ? err[c] : (double)berr[c];
            else
                err[c] += (double)berr[c];
            tsum[c] += (double)bt[c];
        }
    }
    if (tid == 0 && blockIdx.x == 0 && P.tail_queue) P.tail_queue[0] = 0u;   // tail queue of the next gather
    // per-column grid reduction: warp, CTA, one atomic per CTA and column; the last CTA plays ConvergenceManager
#pragma unroll
    for (int c = 0; c < PB; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double e2 = __shfl_xor_sync(0xffffffffu, err[c], o);
            err[c] = is_max ? (err[c] > e2 ? err[c] : e2) : err[c] + e2;
            tsum[c] += __shfl_xor_sync(0xffffffffu, tsum[c], o);
        }
        if (lane == 0) {
            s_red[0][warp][c] = err[c];
            s_red[1][warp][c] = tsum[c];
        }
    }
    __syncthreads();
    if (tid < PB) {
        double e = 0.0, t = 0.0;
        for (int wv = 0; wv < UPP_BLOCK / 32; ++wv) {
            const double e2 = s_red[0][wv][tid];
            e = is_max ? (e > e2 ? e : e2) : e + e2;
            t += s_red[1][wv][tid];
        }
        if (is_max)
            atomicMax((unsigned long long *)&P.sf[tid * PGB_STATE_F64_LEN + PGB_SF_EACC], (unsigned long long)__double_as_longlong(e));
        else
            atomicAdd(&P.sf[tid * PGB_STATE_F64_LEN + PGB_SF_EACC], e);
        atomicAdd(&P.sf[tid * PGB_STATE_F64_LEN + PGB_SF_TACC], t);
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) {
        const int ticket = atomicAdd(&P.si[PB * PGB_STATE_I32_LEN], 1);
        s_last = (ticket == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        if (tid < PB)
            finalize_column(P.sf + tid * PGB_STATE_F64_LEN, P.si + tid * PGB_STATE_I32_LEN,
                            P.err_hist ? P.err_hist + (int64_t)tid * P.hist_stride : nullptr);
        __syncthreads();
        if (tid == 0) {
            volatile int32_t *vsi = P.si;
            bool running = false;
            for (int c = 0; c < PB; ++c) running = running || vsi[c * PGB_STATE_I32_LEN + PGB_SI_STOP] == PGB_RUNNING;
            vsi[PB * PGB_STATE_I32_LEN + 1] = running ? PGB_RUNNING : PGB_CONVERGED;   // the gather's stop word
            vsi[PB * PGB_STATE_I32_LEN + 2] = vsi[PB * PGB_STATE_I32_LEN + 2] + 1;     // steps the panel has executed
            vsi[PB * PGB_STATE_I32_LEN] = 0;
            __threadfence();
        }
    }
}

// Linear texture objects over gather vectors, keyed by (device, address, bytes): the filters alternate two z
// buffers per solve and torch's allocator hands the same blocks out again, so a small LRU table suffices.
// Returns 0 when the vector cannot be a linear texture (alignment, size): the caller then gathers with LDG.
struct TexEntry {
    const void *p;
    size_t bytes;
    int dev, dtype;
    cudaTextureObject_t tex;
    uint64_t stamp;
};
static TexEntry g_tex[64];
static uint64_t g_tex_stamp = 0;
static std::mutex g_tex_mutex;

static cudaTextureObject_t texture_for(const void *z, size_t elems, int dtype /* ELEM_* */) {
    static int enabled = -1;
    if (enabled < 0) {
        const char *e = getenv("PGB_HSELL_TEX");
        enabled = e ? atoi(e) : 1;
    }
    if (!enabled || !z || elems == 0) return 0;
    const size_t bytes = elems * (dtype == ELEM_F32 ? 4 : (dtype == ELEM_F64 ? 8 : 16));
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(g_tex_mutex);
    int victim = 0;
    for (int i = 0; i < 64; ++i) {
        TexEntry &e = g_tex[i];
        if (e.tex && e.p == z && e.bytes == bytes && e.dev == dev && e.dtype == dtype) {
            e.stamp = ++g_tex_stamp;
            return e.tex;
        }
        if (g_tex[i].stamp < g_tex[victim].stamp) victim = i;
    }
    static PerDeviceInt max_linear;
    int &lim = max_linear.here();
    if (lim == 0 && cudaDeviceGetAttribute(&lim, cudaDevAttrMaxTexture1DLinearWidth, dev) != cudaSuccess) lim = 1 << 27;
    if (elems > (size_t)lim || ((uintptr_t)z & 511u)) return 0;
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = const_cast<void *>(z);
    rd.res.linear.desc = dtype == ELEM_F32     ? cudaCreateChannelDesc<float>()
                         : dtype == ELEM_F64   ? cudaCreateChannelDesc<int2>()
                         : dtype == ELEM_F32X4 ? cudaCreateChannelDesc<float4>()
                                               : cudaCreateChannelDesc<int4>();
    rd.res.linear.sizeInBytes = bytes;
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t t = 0;
    if (cudaCreateTextureObject(&t, &rd, &td, nullptr) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    TexEntry &v = g_tex[victim];
    if (v.tex) cudaDestroyTextureObject(v.tex);
    v = TexEntry{z, bytes, dev, dtype, t, ++g_tex_stamp};
    return t;
}

static inline uint64_t mix64_host(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static double g_drop_p = 0.0;
static uint64_t g_drop_seed = 0;
static uint64_t g_drop_launch = 0;   // gather launches since pgb_hsell_set_dropout: every launch draws a new mask

template <typename T, bool TEX, bool ACCUM, bool DROP, bool WEIGHTED = false>
static int launch_gather_variant(const GatherParams &G, int n_ctas, size_t smem, cudaStream_t st) {
    static PerDeviceInt configured_on;
    int &configured = configured_on.here();
    if (!configured) {
        PGB_CUDA_OK(cudaFuncSetAttribute(hsell_gather_kernel<T, TEX, ACCUM, DROP, WEIGHTED>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, HS_SMEM_LIMIT - 64));
        configured = 1;
    }
    hsell_gather_kernel<T, TEX, ACCUM, DROP, WEIGHTED><<<n_ctas, HS_THREADS, smem, st>>>(G);
    PGB_LAUNCH_OK("hsell_gather_kernel");
    return 0;
}

template <typename T>
static int launch_gather(const pgb_hsell *h, const void *z, void *partials, bool accum, const int32_t *stop,
                         uint32_t *tail_queue, int step, cudaStream_t st) {
    const size_t smem = ((size_t)h->block_cols + 1) * sizeof(T);
    if (smem > (size_t)HS_SMEM_LIMIT - 64) return fail("hsell: block_cols=%d does not fit in shared memory", h->block_cols);
    GatherParams G;
    G.h = *h;
    G.z = z;
    G.partials = partials;
    G.stop = stop;
    G.tail_queue = tail_queue;
    G.tail_warps = g_tail_warps;
    if (sizeof(T) == 16) {   // panel elements: 8 tail warps measured best (5 -> 1.394, 8 -> 1.338, 12 -> 1.37 ms, RMAT-24 fp32x4)
        static int panel_tail_warps = -1;
        if (panel_tail_warps < 0) {
            const char *e = getenv("PGB_HSELL_PANEL_TAIL_WARPS");
            panel_tail_warps = e ? atoi(e) : 8;
            if (panel_tail_warps < 0 || panel_tail_warps > HS_WARPS) panel_tail_warps = 8;
        }
        G.tail_warps = panel_tail_warps;
    }
    static int debug_skip = -1;
    if (debug_skip < 0) {
        const char *e = getenv("PGB_HSELL_DEBUG_SKIP");
        debug_skip = e ? atoi(e) : 0;
    }
    G.debug_skip = debug_skip;
    static int bulk = -1;
    if (bulk < 0) {
        const char *e = getenv("PGB_HSELL_BULK");
        bulk = e ? atoi(e) : 1;
    }
    G.bulk = bulk;
    static int tail_batch = -1;
    if (tail_batch < 0) {
        const char *e = getenv("PGB_HSELL_TAIL_BATCH");
        tail_batch = e ? atoi(e) : 1;
        if (tail_batch < 1) tail_batch = 1;
    }
    G.tail_batch = tail_batch;
    G.ztex = h->n_tail_chunks > 0 ? texture_for(z, (size_t)h->seg_len * (size_t)h->n_segments, Elem<T>::kind) : 0;
    if (accum && !h->piece_slice) return fail("hsell: accumulate mode needs pgb_hsell.piece_slice");
    const bool drop = g_drop_p > 0.0;
    (void)step;
    G.drop.key = mix64_host(g_drop_seed ^ ((drop ? ++g_drop_launch : 0) * 0x9E3779B97F4A7C15ull));
    G.drop.thr = drop ? (uint32_t)(g_drop_p * 4294967296.0 > 4294967295.0 ? 4294967295.0 : g_drop_p * 4294967296.0) : 0u;
    G.drop.scale = drop ? (float)(1.0 / (1.0 - g_drop_p)) : 1.0f;
    const int n = h->n_ctas;
    if constexpr (sizeof(T) == 16) {   // panel elements: accumulate mode, no dropout
        if (!accum || drop) return fail("hsell panel: only the accumulate mode without graph_dropout is built");
        return G.ztex ? launch_gather_variant<T, true, true, false>(G, n, smem, st)
                      : launch_gather_variant<T, false, true, false>(G, n, smem, st);
    } else {
        if (h->hub_vals || h->tail_vals) {   // weighted form: accumulate mode, no in-kernel dropout
            if (!h->hub_vals || !h->tail_vals) return fail("hsell: a weighted form needs both hub_vals and tail_vals");
            if (!accum || drop) return fail("hsell: weighted forms run in accumulate mode without in-kernel dropout");
            return G.ztex ? launch_gather_variant<T, true, true, false, true>(G, n, smem, st)
                          : launch_gather_variant<T, false, true, false, true>(G, n, smem, st);
        }
        const int sel = (G.ztex ? 4 : 0) | (accum ? 2 : 0) | (drop ? 1 : 0);
        switch (sel) {
            case 0: return launch_gather_variant<T, false, false, false>(G, n, smem, st);
            case 1: return launch_gather_variant<T, false, false, true>(G, n, smem, st);
            case 2: return launch_gather_variant<T, false, true, false>(G, n, smem, st);
            case 3: return launch_gather_variant<T, false, true, true>(G, n, smem, st);
            case 4: return launch_gather_variant<T, true, false, false>(G, n, smem, st);
            case 5: return launch_gather_variant<T, true, false, true>(G, n, smem, st);
            case 6: return launch_gather_variant<T, true, true, false>(G, n, smem, st);
            default: return launch_gather_variant<T, true, true, true>(G, n, smem, st);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Slot scheduling of a panel job, on the device.  propagate() of B seed columns keeps the PB slots of the panel busy:
// before every step four small launches look at the slots — a column that stopped is written out (its iteration
// count, stop reason and error history recorded) and the next pending column is loaded into the free slot
// (normalisation, scaled start vector, affine term, state: abstract_filters.py:52-56), while the other columns
// keep iterating.  No host round trip: the host enqueues steps in chunks and only polls the count of finished
// columns.  All four launches return at once when no slot changes hands.
// ---------------------------------------------------------------------------------------------
template <typename S, int PB>
__global__ void panel_plan_kernel(const pgb_panel_job J, double *sf, int32_t *si, const double *err_hist) {
    const int s = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (s >= PB) return;
    double *sfs = sf + s * PGB_STATE_F64_LEN;
    int32_t *sis = si + s * PGB_STATE_I32_LEN;
    int col = J.slot_col[s];
    int harvest = -1, load = -1;
    if (col >= 0 && sis[PGB_SI_STOP] != PGB_RUNNING) {
        harvest = col;
        const int steps = sis[PGB_SI_STEPS];
        if (J.col_err)
            for (int k = 1 + lane; k <= steps && k < J.hist_stride; k += 32)
                J.col_err[(int64_t)col * J.hist_stride + k] = err_hist[(int64_t)s * J.hist_stride + k];
        if (lane == 0) {
            int32_t *r = J.col_result + 4 * (int64_t)col;
            r[0] = sis[PGB_SI_ITERATION];
            r[1] = sis[PGB_SI_STOP];
            r[2] = steps;
            r[3] = sfs[PGB_SF_NORM] > 0.0 ? 1 : 0;
            J.plan_norm[s] = sfs[PGB_SF_NORM];
            J.slot_col[s] = -1;
            atomicAdd(&J.sched[1], 1);
        }
        col = -1;
    }
    __syncwarp();
    if (col < 0) {
        if (lane == 0) {
            const int j = J.sched[0] < J.n_cols ? atomicAdd(&J.sched[0], 1) : J.n_cols;
            if (j < J.n_cols) load = j;
        }
        load = __shfl_sync(0xffffffffu, load, 0);
        if (load >= 0) {   // fresh state rows for the new column
            if (lane < PGB_STATE_F64_LEN) sfs[lane] = 0.0;
            if (lane < PGB_STATE_I32_LEN) sis[lane] = (lane == PGB_SI_STOP) ? PGB_CONVERGED : 0;
        }
    }
    if (lane == 0) {
        J.slot_plan[2 * s] = harvest;
        J.slot_plan[2 * s + 1] = load;
    }
}

template <typename S, int PB>
__global__ void __launch_bounds__(256) panel_move_kernel(const pgb_panel_job J, int64_t n, const S *__restrict__ cur_z,
                                                         const S *__restrict__ sq, double *sf) {
    // what a finished column leaves with: its iterate z * sq (affine recursion) or its accumulated result (polynomial)
    const S *__restrict__ cur = J.poly ? (const S *)J.ranks : cur_z;
    __shared__ double scratch[32];
    int harvest[PB], load[PB];
    bool any = false;
#pragma unroll
    for (int s = 0; s < PB; ++s) {
        harvest[s] = J.slot_plan[2 * s];
        load[s] = J.slot_plan[2 * s + 1];
        any = any || harvest[s] >= 0 || load[s] >= 0;
    }
    if (!any) return;
    const S *__restrict__ cols = (const S *)J.cols;
    S *__restrict__ out = (S *)J.out;
    double scale[PB], norm[PB];
#pragma unroll
    for (int s = 0; s < PB; ++s) {
        scale[s] = (harvest[s] >= 0 && J.preserve_norm && J.plan_norm[s] > 0.0) ? J.plan_norm[s] : 1.0;
        norm[s] = 0.0;
    }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t user = J.perm ? (int64_t)J.perm[i] : i;
        bool out_any = false;
#pragma unroll
        for (int s = 0; s < PB; ++s) out_any = out_any || harvest[s] >= 0;
        if (out_any) {
            const Pack<S, PB> z = ld_pack_rw(cur + i * PB);
            const S sqi = J.poly ? (S)1 : sq[i];
#pragma unroll
            for (int s = 0; s < PB; ++s)
                if (harvest[s] >= 0) {
                    S v = z.v[s] * sqi;
                    if (scale[s] != 1.0) v = (S)(v * (S)scale[s]);
                    out[user * J.out_row_stride + harvest[s] * J.out_col_stride] = v;
                }
        }
#pragma unroll
        for (int s = 0; s < PB; ++s)
            if (load[s] >= 0) norm[s] += fabs((double)cols[user * J.row_stride + load[s] * J.col_stride]);
    }
#pragma unroll
    for (int s = 0; s < PB; ++s)
        if (load[s] >= 0) {   // uniform across the grid
            const double t = block_sum(norm[s], scratch);
            if (threadIdx.x == 0) atomicAdd(&sf[s * PGB_STATE_F64_LEN + PGB_SF_NORM], t);
        }
}

template <typename S, int PB>
__global__ void __launch_bounds__(256) panel_load_kernel(const pgb_panel_job J, int64_t n, S *__restrict__ cur,
                                                         S *__restrict__ q, const S *__restrict__ sq,
                                                         const S *__restrict__ c, double *sf) {
    __shared__ double scratch[32];
    int load[PB];
    bool any = false;
#pragma unroll
    for (int s = 0; s < PB; ++s) {
        load[s] = J.slot_plan[2 * s + 1];
        any = any || load[s] >= 0;
    }
    if (!any) return;
    const S *__restrict__ cols = (const S *)J.cols;
    const S *__restrict__ coefvec = (const S *)J.coefvec;
    double norm[PB], coef[PB], tacc[PB], bacc[PB];
#pragma unroll
    for (int s = 0; s < PB; ++s) {
        norm[s] = load[s] >= 0 ? sf[s * PGB_STATE_F64_LEN + PGB_SF_NORM] : 0.0;
        coef[s] = (load[s] >= 0 && J.col_params) ? J.col_params[3 * (int64_t)load[s] + 2] : J.coef;
        tacc[s] = bacc[s] = 0.0;
    }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t user = J.perm ? (int64_t)J.perm[i] : i;
        const S sqi = sq[i], ci = J.poly ? (S)0 : c[i];
        const double cv = coefvec ? (double)coefvec[i] : 0.0;
#pragma unroll
        for (int s = 0; s < PB; ++s)
            if (load[s] >= 0) {
                S zi = (S)0, qi = (S)0;
                if (norm[s] > 0.0) {                                     // abstract_filters.py:53-55
                    const S pn = (S)((double)cols[user * J.row_stride + load[s] * J.col_stride] / norm[s]);
                    zi = pn / sqi;
                    if (!J.poly) {
                        qi = (S)((coefvec ? cv : coef[s]) * (double)pn) / sqi;
                        tacc[s] += (double)zi * (double)ci;
                        bacc[s] += (double)qi * (double)sqi;
                    }
                }
                cur[i * PB + s] = zi;
                if (J.poly)
                    ((S *)J.ranks)[i * PB + s] = (S)0;                   // abstract_filters.py:213
                else
                    q[i * PB + s] = qi;
            }
    }
#pragma unroll
    for (int s = 0; s < PB; ++s)
        if (load[s] >= 0) {
            const double t = block_sum(tacc[s], scratch);
            const double b = block_sum(bacc[s], scratch);
            if (threadIdx.x == 0) {
                atomicAdd(&sf[s * PGB_STATE_F64_LEN + PGB_SF_TACC], t);
                atomicAdd(&sf[s * PGB_STATE_F64_LEN + PGB_SF_BIAS], b);
            }
        }
}

template <int PB>
__global__ void panel_finish_kernel(const pgb_panel_job J, double *sf, int32_t *si) {
    const int s = threadIdx.x;
    if (s < PB) {
        const int load = J.slot_plan[2 * s + 1];
        if (load >= 0) {
            double *sfs = sf + s * PGB_STATE_F64_LEN;
            int32_t *sis = si + s * PGB_STATE_I32_LEN;
            const double amul = J.col_params ? J.col_params[3 * (int64_t)load] : J.alpha;
            const double alpha_s = J.col_params ? J.col_params[3 * (int64_t)load + 1] : J.alpha_s;
            const double t = sfs[PGB_SF_TACC];
            const bool live = sfs[PGB_SF_NORM] > 0.0;
            sfs[PGB_SF_TACC] = 0.0;
            sfs[PGB_SF_EACC] = 0.0;
            sfs[PGB_SF_ALPHA] = alpha_s;
            sfs[PGB_SF_AMUL] = amul;
            sfs[PGB_SF_TOL] = J.tol;
            sfs[PGB_SF_MEAN] = J.mean;
            sfs[PGB_SF_INVS] = (J.quotient && live) ? 1.0 / (alpha_s * t + sfs[PGB_SF_BIAS]) : 1.0;
            sis[PGB_SI_STEPS] = 0;
            sis[PGB_SI_ITERATION] = 0;
            sis[PGB_SI_MAX_ITERS] = J.max_iters;
            sis[PGB_SI_END_MODULO] = J.end_modulo;
            sis[PGB_SI_ERR_MODE] = J.err_mode;
            sis[PGB_SI_QUOTIENT] = J.quotient;
            sis[PGB_SI_STOP] = live ? PGB_RUNNING : PGB_CONVERGED;   // a zero personalization never iterates (:53-54)
            J.slot_col[s] = load;
        }
    }
    __syncthreads();
    if (s == 0) {
        bool running = false;
        for (int k = 0; k < PB; ++k) running = running || si[k * PGB_STATE_I32_LEN + PGB_SI_STOP] == PGB_RUNNING;
        si[PB * PGB_STATE_I32_LEN + 1] = running ? PGB_RUNNING : PGB_CONVERGED;   // the gather's stop word
    }
}

template <typename S, typename V, int PB>
static int panel_steps(const pgb_hsell *h, const PanelStep &P0, const pgb_panel_job &J, void *zbuf0, void *zbuf1,
                       int first_step, int num_launches, bool symdeg, cudaStream_t st) {
    PanelStep P = P0;
    void *buf[2] = {zbuf0, zbuf1};
    const int32_t *stop = P.si + PB * PGB_STATE_I32_LEN + 1;
    static PerDeviceInt ctas_on[2];
    int &ctas = ctas_on[symdeg ? 1 : 0].here();
    if (ctas == 0) {
        int v = 0;
        const cudaError_t e = symdeg
            ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, hsell_update_panel_kernel<S, PB, true, MODE_AFFINE>, UPP_BLOCK, 0)
            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, hsell_update_panel_kernel<S, PB, false, MODE_AFFINE>, UPP_BLOCK, 0);
        ctas = (e == cudaSuccess && v >= 1) ? v : 4;
    }
    const bool poly = J.poly != 0;
    int64_t want = ceil_div(P.n, (int64_t)UPP_BLOCK * UPP_ROWS);
    const int64_t cap = (int64_t)sm_count() * ctas;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    for (int j = 0; j < num_launches; ++j) {
        const int k = first_step + j;
        P.zin = buf[(k - 1) & 1];
        P.zout = buf[k & 1];
        // slots change hands on the buffer this step reads
        const int mgrid = stride_grid(P.n, 256) < 1184 ? stride_grid(P.n, 256) : 1184;
        panel_plan_kernel<S, PB><<<1, 32 * PB, 0, st>>>(J, P.sf, P.si, P.err_hist);
        panel_move_kernel<S, PB><<<mgrid, 256, 0, st>>>(J, P.n, (const S *)P.zin, (const S *)J.sq, P.sf);
        panel_load_kernel<S, PB><<<mgrid, 256, 0, st>>>(J, P.n, (S *)const_cast<void *>(P.zin), (S *)const_cast<void *>(P.q),
                                                        (const S *)J.sq, (const S *)P.c, P.sf);
        panel_finish_kernel<PB><<<1, 32, 0, st>>>(J, P.sf, P.si);
        PGB_LAUNCH_OK("panel scheduling kernels");
        if (launch_gather<V>(h, P.zin, P.y, true, stop, P.tail_queue, k, st)) return 1;
        if (symdeg && !poly)
            hsell_update_panel_kernel<S, PB, true, MODE_AFFINE><<<(int)want, UPP_BLOCK, 0, st>>>(P);
        else if (!poly)
            hsell_update_panel_kernel<S, PB, false, MODE_AFFINE><<<(int)want, UPP_BLOCK, 0, st>>>(P);
        else if (symdeg)
            hsell_update_panel_kernel<S, PB, true, MODE_POLY><<<(int)want, UPP_BLOCK, 0, st>>>(P);
        else
            hsell_update_panel_kernel<S, PB, false, MODE_POLY><<<(int)want, UPP_BLOCK, 0, st>>>(P);
        PGB_LAUNCH_OK("hsell_update_panel_kernel");
    }
    return 0;
}

template <typename T>
static int launch_reduce(const pgb_hsell *h, void *partials, const int32_t *stop, cudaStream_t st) {
    if (h->n_reduce <= 0) return 0;
    int64_t want = ceil_div(h->n_reduce, 8);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (want > cap) want = cap;
    hsell_reduce_kernel<T><<<(int)want, 256, 0, st>>>(*h, (T *)partials, stop);
    PGB_LAUNCH_OK("hsell_reduce_kernel");
    return 0;
}

template <typename T, int MODE, bool SYMDEG, int GROUP>
static int launch_update_g(const StepParams &P, const pgb_hsell *h, const void *partials, cudaStream_t st) {
    int64_t want = ceil_div(h->n_slices, UPD_WARPS * GROUP);
    if (want < h->n_heavy) want = h->n_heavy;
    static PerDeviceInt ctas_on;
    int &ctas = ctas_on.here();
    if (ctas == 0) {
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, hsell_update_kernel<T, MODE, SYMDEG, GROUP>, UPD_BLOCK, 0) != cudaSuccess || v < 1)
            v = 2;
        ctas = v;
    }
    const int64_t cap = (int64_t)sm_count() * ctas;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    hsell_update_kernel<T, MODE, SYMDEG, GROUP><<<(int)want, UPD_BLOCK, 0, st>>>(P, *h, (const T *)partials);
    PGB_LAUNCH_OK("hsell_update_kernel");
    return 0;
}

// slices a warp of the update kernel handles together: 4 (fewer, fatter warps) or 2 (PGB_HSELL_UPD_GROUP)
template <typename T, int MODE, bool SYMDEG>
static int launch_update(const StepParams &P, const pgb_hsell *h, const void *partials, cudaStream_t st) {
    static int group = 0;
    if (group == 0) {
        const char *e = getenv("PGB_HSELL_UPD_GROUP");
        group = (e && atoi(e) == 2) ? 2 : 4;
    }
    return group == 2 ? launch_update_g<T, MODE, SYMDEG, 2>(P, h, partials, st)
                      : launch_update_g<T, MODE, SYMDEG, 4>(P, h, partials, st);
}

template <typename T, int MODE, bool SYMDEG>
static int launch_update_accum(const StepParams &P, void *yacc, cudaStream_t st) {
    static PerDeviceInt ctas_on;
    int &ctas = ctas_on.here();
    if (ctas == 0) {
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, hsell_update_accum_kernel<T, MODE, SYMDEG>, UPA_BLOCK, 0) != cudaSuccess || v < 1)
            v = 4;
        ctas = v;
    }
    int64_t want = ceil_div(P.n, (int64_t)UPA_BLOCK * UPA_ROWS);
    const int64_t cap = (int64_t)sm_count() * ctas;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    hsell_update_accum_kernel<T, MODE, SYMDEG><<<(int)want, UPA_BLOCK, 0, st>>>(P, (T *)yacc);
    PGB_LAUNCH_OK("hsell_update_accum_kernel");
    return 0;
}

template <typename T, int MODE>
static int hsell_step_accum(const StepParams &P, const pgb_hsell *h, bool symdeg, const int32_t *stop, cudaStream_t st) {
    if (launch_gather<T>(h, P.zin, P.yacc, true, stop, P.span_cnt, P.step, st)) return 1;
    return symdeg ? launch_update_accum<T, MODE, true>(P, P.yacc, st) : launch_update_accum<T, MODE, false>(P, P.yacc, st);
}

// One fused step on the hsell form: gather (kernel A) + update (kernel B).  Called from spmv_fused.cu.
template <int MODE>
int hsell_step(const StepParams &P, const pgb_hsell *h, void *partials, int dtype, bool symdeg, cudaStream_t st) {
    if (P.yacc) {   // accumulate mode: pieces are RED.ADDed into y, the update pass streams y (two launches per step)
        if (!P.span_cnt) return fail("hsell: the workspace has no counter array");
        if (h->n_rows != P.n) return fail("hsell: form built for %lld rows, graph has %lld", (long long)h->n_rows, (long long)P.n);
        const int32_t *stop_a = (MODE == MODE_CONV) ? nullptr : P.si + PGB_SI_STOP;
        if (dtype == PGB_F32) return hsell_step_accum<float, MODE>(P, h, symdeg, stop_a, st);
        if (dtype == PGB_F64) return hsell_step_accum<double, MODE>(P, h, symdeg, stop_a, st);
        return fail("unknown dtype %d", dtype);
    }
    if (!partials) return fail("hsell: the workspace has no partials buffer");
    if (!P.span_cnt) return fail("hsell: the workspace has no counter array");
    if (h->n_rows != P.n) return fail("hsell: form built for %lld rows, graph has %lld", (long long)h->n_rows, (long long)P.n);
    const int32_t *stop = (MODE == MODE_CONV) ? nullptr : P.si + PGB_SI_STOP;
    if (dtype == PGB_F32) {
        if (launch_gather<float>(h, P.zin, partials, false, stop, P.span_cnt, P.step, st)) return 1;
        if (launch_reduce<float>(h, partials, stop, st)) return 1;
        return symdeg ? launch_update<float, MODE, true>(P, h, partials, st)
                      : launch_update<float, MODE, false>(P, h, partials, st);
    } else if (dtype == PGB_F64) {
        if (launch_gather<double>(h, P.zin, partials, false, stop, P.span_cnt, P.step, st)) return 1;
        if (launch_reduce<double>(h, partials, stop, st)) return 1;
        return symdeg ? launch_update<double, MODE, true>(P, h, partials, st)
                      : launch_update<double, MODE, false>(P, h, partials, st);
    }
    return fail("unknown dtype %d", dtype);
}

template int hsell_step<MODE_CONV>(const StepParams &, const pgb_hsell *, void *, int, bool, cudaStream_t);
template int hsell_step<MODE_AFFINE>(const StepParams &, const pgb_hsell *, void *, int, bool, cudaStream_t);
template int hsell_step<MODE_POLY>(const StepParams &, const pgb_hsell *, void *, int, bool, cudaStream_t);

}  // namespace pgb

using namespace pgb;

extern "C" {

int pgb_hsell_max_block_cols(int dtype) {
    const int bytes = dtype == PGB_F32 ? 4 : (dtype == PGB_F64 ? 8 : 0);
    if (!bytes) return 0;
    int cols = (HS_SMEM_LIMIT - 64) / bytes - 1;
    if (cols > 65535) cols = 65535;
    return cols & ~63;   // keeps every block start 16-byte aligned for the vector loader
}

int pgb_hsell_panel_width(int dtype) { return dtype == PGB_F32 ? 4 : (dtype == PGB_F64 ? 2 : 0); }

int pgb_hsell_panel_block_cols(void) { return (128 * 1024) / 16; }

int pgb_affine_steps_panel(const pgb_hsell *h, const int32_t *indptr, int dtype, const pgb_panel_job *job, const void *w,
                           const void *c, void *q, void *zbuf0, void *zbuf1, double *state_f64, int32_t *state_i32,
                           double *err_hist, void *yacc, uint32_t *tail_queue, int first_step, int num_launches,
                           void *stream) {
    if (!h) return fail("pgb_affine_steps_panel: null hsell form");
    if (!job) return fail("pgb_affine_steps_panel: null job");
    if (h->n_segments != 1) return fail("pgb_affine_steps_panel: row-partitioned forms are not supported");
    if (!h->piece_slice) return fail("pgb_affine_steps_panel needs pgb_hsell.piece_slice");
    if (!w && !indptr) return fail("pgb_affine_steps_panel: degree-derived factors need the row pointers");
    if (first_step < 1) return fail("pgb_affine_steps_panel: first_step must be >= 1");
    if (!yacc || !tail_queue || !state_f64 || !state_i32 || !err_hist)
        return fail("pgb_affine_steps_panel: the state arrays, err_hist, yacc and tail_queue are required");
    if (!job->poly && (!q || !c)) return fail("pgb_affine_steps_panel: the affine recursion needs q and c");
    if (job->poly && (!job->ranks || !job->coef_table)) return fail("pgb_affine_steps_panel: a polynomial job needs ranks and coef_table");
    if (job->poly && job->quotient) return fail("pgb_affine_steps_panel: polynomial filters have no quotient");
    if (!job->cols || !job->out || !job->sq || !job->sched || !job->slot_col || !job->slot_plan || !job->plan_norm ||
        !job->col_result)
        return fail("pgb_affine_steps_panel: incomplete job (cols, out, sq, sched, slot_col, slot_plan, plan_norm, col_result)");
    if (job->n_cols < 0 || job->hist_stride < 2) return fail("pgb_affine_steps_panel: bad n_cols / hist_stride");
    if (h->n_rows == 0 || job->n_cols == 0) return 0;
    const int PB = pgb_hsell_panel_width(dtype);
    if (!PB) return fail("pgb_affine_steps_panel: unknown dtype %d", dtype);
    PanelStep P;
    memset(&P, 0, sizeof(P));
    P.n = h->n_rows;
    P.indptr = indptr;
    P.q = q;
    P.w = w;
    P.sq = w ? job->sq : nullptr;
    P.c = c;
    P.y = yacc;
    P.sf = state_f64;
    P.si = state_i32;
    P.err_hist = err_hist;
    P.hist_stride = job->hist_stride;
    P.tail_queue = tail_queue;
    P.ranks = job->ranks;
    P.coef = job->coef_table;
    const bool symdeg = (w == nullptr);
    if (dtype == PGB_F32)
        return panel_steps<float, f32x4, 4>(h, P, *job, zbuf0, zbuf1, first_step, num_launches, symdeg, as_stream(stream));
    return panel_steps<double, f64x2, 2>(h, P, *job, zbuf0, zbuf1, first_step, num_launches, symdeg, as_stream(stream));
}

int pgb_hsell_set_dropout(double p, uint64_t seed) {
    if (!(p >= 0.0) || p >= 1.0) return fail("pgb_hsell_set_dropout: p must be in [0, 1)");
    g_drop_p = p;
    g_drop_seed = seed;
    g_drop_launch = 0;
    return 0;
}

int pgb_hsell_set_tail_warps(int warps) {
    if (warps < 0 || warps > HS_WARPS) return fail("pgb_hsell_set_tail_warps: %d is not in 0..%d", warps, HS_WARPS);
    g_tail_warps = warps;
    return 0;
}

int pgb_hsell_count(int64_t n, const int32_t *indptr, const int32_t *indices, int32_t block_cols, int32_t n_blocks,
                    int32_t min_entries, double round_cost, int32_t n_segments, int64_t seg_len, int32_t n_windows,
                    int64_t window_len, int32_t window_min_rounds, int32_t *hub_rounds, int32_t *tail_rounds,
                    void *stream) {
    if (n <= 0) return 0;
    if (block_cols < 1 || block_cols > 65535) return fail("pgb_hsell_count: block_cols must be in 1..65535");
    if (n_blocks < 0 || (int64_t)n_blocks * block_cols >= (1ll << 31)) return fail("pgb_hsell_count: bad n_blocks");
    if (n_windows < 1 || n_windows > HS_MAX_WINDOWS || (n_windows > 1 && window_len < 1))
        return fail("pgb_hsell_count: n_windows must be in 1..%d with a positive window_len", HS_MAX_WINDOWS);
    if (n_segments < 1 || block_cols % n_segments) return fail("pgb_hsell_count: block_cols must be a multiple of n_segments");
    const int64_t n_slices = ceil_div(n, 32);
    const int grid = stride_grid(n_slices * 32, 256);
    hsell_count_kernel<<<grid, 256, 0, as_stream(stream)>>>(n, n_slices, indptr, indices, block_cols, n_blocks,
                                                            min_entries, (float)round_cost, n_segments, seg_len,
                                                            n_windows, window_len, window_min_rounds, hub_rounds,
                                                            tail_rounds);
    PGB_LAUNCH_OK("hsell_count_kernel");
    return 0;
}

int pgb_hsell_fill(int64_t n, const int32_t *indptr, const int32_t *indices, int32_t block_cols, int32_t n_blocks,
                   int32_t n_segments, int64_t seg_len, const int32_t *hub_rounds, const int32_t *tail_rounds,
                   const int64_t *hub_round_base, const int64_t *hub_part_base, const int64_t *tail_round_base,
                   const int64_t *tail_part_base, const int32_t *slice_ptr, uint32_t *hub_words, int32_t *tail_cols,
                   int32_t *piece_row, int32_t *scratch, int32_t banks, int32_t n_windows, int64_t window_len,
                   int dtype, const void *values, void *hub_vals, void *tail_vals, void *stream) {
    if (n <= 0) return 0;
    if (scratch && (banks < 1 || banks > 32 || (banks & (banks - 1)))) return fail("pgb_hsell_fill: banks must be a power of two <= 32");
    if (scratch && block_cols >= 0xffff) return fail("pgb_hsell_fill: bank-aware ordering needs block_cols < 65535");
    if (n_segments < 1 || block_cols % n_segments) return fail("pgb_hsell_fill: block_cols must be a multiple of n_segments");
    if (n_windows < 1 || n_windows > HS_MAX_WINDOWS || (n_windows > 1 && window_len < 1))
        return fail("pgb_hsell_fill: n_windows must be in 1..%d with a positive window_len", HS_MAX_WINDOWS);
    if (values && (!hub_vals || !tail_vals)) return fail("pgb_hsell_fill: a weighted form needs hub_vals and tail_vals");
    if (values && dtype != PGB_F32 && dtype != PGB_F64) return fail("pgb_hsell_fill: unknown dtype %d", dtype);
    const int64_t n_slices = ceil_div(n, 32);
    const int grid = stride_grid(n_slices * 32 * ((n_blocks + FILL_GROUP - 1) / FILL_GROUP + 1), 256);
    cudaStream_t st = as_stream(stream);
#define PGB_FILL_ARGS(VT)                                                                                           \
    n, n_slices, indptr, indices, (const VT *)values, (VT *)hub_vals, (VT *)tail_vals, block_cols, n_blocks,       \
        n_segments, seg_len, hub_rounds, tail_rounds, hub_round_base, hub_part_base, tail_round_base,              \
        tail_part_base, slice_ptr, hub_words, tail_cols, piece_row, scratch, banks, n_windows, window_len
    if (!values)
        hsell_fill_kernel<float, false><<<grid, 256, 0, st>>>(PGB_FILL_ARGS(float));
    else if (dtype == PGB_F32)
        hsell_fill_kernel<float, true><<<grid, 256, 0, st>>>(PGB_FILL_ARGS(float));
    else
        hsell_fill_kernel<double, true><<<grid, 256, 0, st>>>(PGB_FILL_ARGS(double));
#undef PGB_FILL_ARGS
    PGB_LAUNCH_OK("hsell_fill_kernel");
    return 0;
}

}  // extern "C"
