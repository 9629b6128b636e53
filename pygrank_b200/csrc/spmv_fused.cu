// K2 / K4 / K5 — the merge-path per-iteration kernel (weighted graphs; A/B twin of csrc/hsell.cu on
// unweighted ones) and the C-ABI entry points of the fused steps, which dispatch to either kernel family.
//
// Reference op sequence replaced per iteration (paths under /root/reference/pygrank):
//   conv(ranks, M)                      core/backend/numpy.py:64-65 (scipy csc_matvec)
//   * alpha + personalization*(1-alpha) algorithms/filters/adhoc.py:36       (3 vector passes)
//   safe_div(ranks, sum(ranks))         algorithms/filters/abstract_filters.py:133-134 (2 passes)
//   Mabs(prev)(cur) <= tol              algorithms/convergence.py:96-101, measures/supervised.py:101-106
//
// Work decomposition.  The CSR is cut by merge path: the sequence "entries of row 0, end-marker of
// row 0, entries of row 1, end-marker of row 1, ..." (n + nnz items, stored as the item stream) is split
// into tiles of TILE_ITEMS consecutive items, so every tile costs the same no matter how skewed the
// degrees are (power-law hubs, 38 % empty rows on RMAT).  A persistent grid (CTAs-per-SM x 148 SMs) of
// warp-autonomous tiles walks them (item_stream_kernel below).  Rows cut by a tile boundary are completed
// by the LAST tile that reaches them (fp64 atomic partial + ticket per completing tile; no spinning, no
// ordering assumption).  Grid-level sums (error numerator, next normaliser) are one fp64 atomic per CTA;
// the last CTA to finish plays ConvergenceManager on the device (step_common.cuh).
#include <stdlib.h>

#include "common.cuh"
#include "step_common.cuh"

namespace pgb {

constexpr int BLOCK = 256;
#ifndef PGB_IPT
#define PGB_IPT 9
#endif
constexpr int IPT = PGB_IPT;  // odd: thread-blocked reads of shared memory hit distinct banks
static_assert(IPT % 2 == 1 && IPT <= 31, "items per thread must be odd");
constexpr int TILE_ITEMS = BLOCK * IPT;   // merge items per tile (one warp walks a tile in 8 passes)
constexpr int WARPS = BLOCK / 32;
constexpr int SUB_ITEMS = 32 * IPT;       // items one warp consumes per pass
static_assert(TILE_ITEMS % SUB_ITEMS == 0, "a tile is a whole number of warp passes");

__global__ void state_finalize_kernel(double *sf, int32_t *si, double *err_hist) {
    if (si[PGB_SI_STOP] != PGB_RUNNING) return;
    finalize_state(sf, si, err_hist);
}

// Cross-tile row completion without a gpu-scope fence.  __threadfence() compiles to
// MEMBAR.SC.GPU + CCTL.IVALL, and the CCTL invalidates the whole L1 of the SM — issued once or twice
// per tile by every warp it kept the gather vector out of L1 (6 % hit rate measured).  Both atomics
// below are performed at L2; the ticket increment is made data-dependent on the RETURN of the partial
// sum's atomic, so it cannot be issued before that add has been performed, and the last arrival reads
// the total with another L2 atomic after it has seen every ticket.
__device__ __forceinline__ bool span_arrive(double *acc, uint32_t *cnt, double partial, uint32_t expected,
                                            double *total) {
    const double old = atomicAdd(acc, partial);
    // never true for finite data; forces the scoreboard wait on `old` before the ticket is issued
    const uint32_t inc = (__double_as_longlong(old) == 0x7ff8dead00000001ll) ? 2u : 1u;
    const uint32_t arrived = atomicAdd(cnt, inc);
    if (arrived != expected - 1) return false;
    const unsigned long long raw = atomicExch((unsigned long long *)acc, 0ull);
    atomicExch(cnt, 0u);
    *total = __longlong_as_double((long long)raw);
    return true;
}

// ---------------------------------------------------------------------------------------------
// Warp-autonomous tiles over the ITEM-SPACE stream.  The graph is additionally stored as
// istream[n + nnz]: for every row its ascending column indices followed by one terminator
// -1-deg (same bytes as indices + row pointers).  A merge tile is then a plain slice of that array:
// no row-pointer loads, no rank arithmetic — a lane's item is an entry (col >= 0) or a row end
// (col < 0), and the terminator carries the degree the SYMDEG update needs.  The dependent chain
// of a pass is item load -> gather -> merge, and it is software-pipelined: the items of pass i+1
// are loaded, and the finished rows of pass i-1 are updated, while the gathers of pass i fly.
#ifndef PGB_V3_MINB
#define PGB_V3_MINB 3   // 80 registers/thread: the pipelined pass keeps 3 x IPT values live without spilling
#endif
template <typename T, bool WEIGHTED, int MODE, bool SYMDEG>
__global__ void __launch_bounds__(BLOCK, PGB_V3_MINB) item_stream_kernel(const StepParams P) {
    // one value buffer per warp: the row sums of pass i-1 are consumed (step c) before the items of
    // pass i are parked in it (step d); only the small degree buffer is double buffered.  Shared
    // memory is kept small on purpose: what it does not take stays L1 for the gathers.
    __shared__ T s_buf[WARPS][SUB_ITEMS + 1];
    __shared__ uint16_t s_deg[WARPS][2][SYMDEG ? SUB_ITEMS + 2 : 2];  // row degrees (0xFFFF: look it up)
    __shared__ unsigned s_mask[WARPS][IPT + 1];
    __shared__ int s_pre[WARPS][IPT + 1];
    __shared__ double s_red[32];

    if (MODE != MODE_CONV) {
        if (P.si[PGB_SI_STOP] != PGB_RUNNING) return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int32_t PAD = (int32_t)0x80000000;  // items past the end of the stream
    const T *__restrict__ zin = (const T *)P.zin;
    const int32_t *__restrict__ istream = P.istream;
    const T *__restrict__ vstream = (const T *)P.vstream;
    if (lane == 0) s_mask[warp][IPT] = 0u;
    RowUpdate<T, MODE, SYMDEG> update(P);
    const int64_t total_items = P.n + P.nnz;

    for (int32_t tile = blockIdx.x * WARPS + warp; tile < P.n_tiles; tile += gridDim.x * WARPS) {
        const int64_t item_lo = (int64_t)tile * TILE_ITEMS;
        const int64_t item_hi = (item_lo + TILE_ITEMS < total_items) ? item_lo + TILE_ITEMS : total_items;
        const int32_t r_lo = P.tile_row[tile], r_hi = P.tile_row[tile + 1];
        // the first row finished here began in an earlier tile iff its first item precedes the tile
        const bool lead_span = (r_hi > r_lo) && ((int64_t)P.indptr[r_lo] + r_lo < item_lo);
        int64_t r_cur = r_lo;
        T carry = (T)0;
        int cur = 0;
        int pend_rows = 0, pend_buf = 0;   // finished rows of the previous pass, not yet updated
        int64_t pend_r = 0;

        int32_t nxt[IPT];
        T nxtv[WEIGHTED ? IPT : 1];
#pragma unroll
        for (int s = 0; s < IPT; ++s) {
            const int64_t g = item_lo + s * 32 + lane;
            nxt[s] = (g < item_hi) ? ld_stream(istream + g) : PAD;
            if (WEIGHTED) nxtv[s] = (g < item_hi) ? ld_stream(vstream + g) : (T)0;
        }

        for (int64_t I0 = item_lo; I0 < item_hi; I0 += SUB_ITEMS) {
            T *buf = s_buf[warp];
            // ---- a. classify the items of this pass, launch the gathers -------------------------
            int32_t cols[IPT];
            T x[IPT];
            unsigned m[IPT];
            int nrows = 0;
#pragma unroll
            for (int s = 0; s < IPT; ++s) {
                cols[s] = nxt[s];
                m[s] = __ballot_sync(FULL, cols[s] < 0 && cols[s] != PAD);
            }
#pragma unroll
            for (int s = 0; s < IPT; ++s) x[s] = (cols[s] >= 0) ? __ldg(zin + cols[s]) : (T)0;
            if (WEIGHTED) {
#pragma unroll
                for (int s = 0; s < IPT; ++s) x[s] *= nxtv[WEIGHTED ? s : 0];
            }
#pragma unroll
            for (int s = 0; s < IPT; ++s) {
                if (lane == s) {
                    s_mask[warp][s] = m[s];
                    s_pre[warp][s] = nrows;
                }
                if (SYMDEG) {
                    if (cols[s] < 0 && cols[s] != PAD) {
                        const int deg = -1 - cols[s];
                        s_deg[warp][cur][nrows + __popc(m[s] & lt_mask)] = (uint16_t)(deg < 0xFFFF ? deg : 0xFFFF);
                    }
                }
                nrows += __popc(m[s]);
            }
            // ---- b. prefetch the items of the next pass -----------------------------------------
            {
                const int64_t nbase = I0 + SUB_ITEMS;
#pragma unroll
                for (int s = 0; s < IPT; ++s) {
                    const int64_t g = nbase + s * 32 + lane;
                    nxt[s] = (g < item_hi) ? ld_stream(istream + g) : PAD;
                    if (WEIGHTED) nxtv[s] = (g < item_hi) ? ld_stream(vstream + g) : (T)0;
                }
            }
            // ---- c. update the rows finished in the previous pass (loads fly with the gathers) ----
            if (pend_rows > 0) {
                const T *rs = s_buf[warp];
                for (int k = lane; k < pend_rows; k += 32) {
                    const int64_t row = pend_r + k;
                    int deg = SYMDEG ? (int)s_deg[warp][pend_buf][k] : 0;
                    if (SYMDEG && deg == 0xFFFF) deg = P.indptr[row + 1] - P.indptr[row];
                    if (lead_span && row == r_lo) {
                        const int64_t t_a = ((int64_t)P.indptr[row] + row) / TILE_ITEMS;
                        const uint32_t expected = (uint32_t)(tile - t_a + 1);
                        double total;
                        if (span_arrive(&P.span_acc[tile], &P.span_cnt[tile], (double)rs[k], expected, &total))
                            update(row, (T)total, deg);
                    } else {
                        update(row, rs[k], deg);
                    }
                }
                pend_rows = 0;
                __syncwarp();  // the value buffer is free again
            }
            // ---- d. park the gathered values in item space, merge blocked -----------------------
            if (nrows == 0) {
                // interior of a long row (about half of all passes on power-law graphs): no row ends,
                // so the pass is one warp-wide sum added to the open row
                T tot = x[0];
#pragma unroll
                for (int s = 1; s < IPT; ++s) tot += x[s];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
                carry += tot;
                continue;
            }
#pragma unroll
            for (int s = 0; s < IPT; ++s) buf[s * 32 + lane] = x[s];
            __syncwarp();
            {
                const int p0 = lane * IPT;
                const int w0 = p0 >> 5, sh = p0 & 31;
                const unsigned lo = s_mask[warp][w0], hi = s_mask[warp][w0 + 1];
                int k = s_pre[warp][w0] + __popc(lo & ((1u << sh) - 1u));
                const unsigned bits = __funnelshift_r(lo, hi, sh) & ((1u << IPT) - 1u);
                T it[IPT];
#pragma unroll
                for (int j = 0; j < IPT; ++j) it[j] = buf[p0 + j];
                __syncwarp();  // items are in registers: the buffer now receives the row sums
                T run = (T)0, head = (T)0;
                int first_k = -1;
#pragma unroll
                for (int j = 0; j < IPT; ++j) {
                    if ((bits >> j) & 1u) {
                        if (first_k < 0) {
                            head = run;
                            first_k = k;
                        } else {
                            buf[k] = run;
                        }
                        run = (T)0;
                        ++k;
                    } else {
                        run += it[j];
                    }
                }
                const bool closed = first_k >= 0;
                const unsigned closed_mask = __ballot_sync(FULL, closed);
                T seg = run;
                bool f = closed;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const T t = __shfl_up_sync(FULL, seg, d);
                    const bool tf = __shfl_up_sync(FULL, (int)f, d);
                    if (lane >= d && !f) {
                        seg += t;
                        f = tf;
                    }
                }
                T carry_in = __shfl_up_sync(FULL, seg, 1);
                if (lane == 0) carry_in = (T)0;
                if ((closed_mask & lt_mask) == 0u) carry_in += carry;
                if (closed) buf[first_k] = head + carry_in;
                const T last = __shfl_sync(FULL, seg, 31);
                carry = (closed_mask == 0u) ? carry + last : last;
            }
            __syncwarp();
            pend_rows = nrows;
            pend_buf = cur;
            pend_r = r_cur;
            r_cur += nrows;
            cur ^= 1;
            if (I0 + SUB_ITEMS >= item_hi && pend_rows > 0) {
                // last pass of the tile: nothing left to hide behind, update now
                const T *rs = s_buf[warp];
                for (int k = lane; k < pend_rows; k += 32) {
                    const int64_t row = pend_r + k;
                    int deg = SYMDEG ? (int)s_deg[warp][pend_buf][k] : 0;
                    if (SYMDEG && deg == 0xFFFF) deg = P.indptr[row + 1] - P.indptr[row];
                    if (lead_span && row == r_lo) {
                        const int64_t t_a = ((int64_t)P.indptr[row] + row) / TILE_ITEMS;
                        const uint32_t expected = (uint32_t)(tile - t_a + 1);
                        double total;
                        if (span_arrive(&P.span_acc[tile], &P.span_cnt[tile], (double)rs[k], expected, &total))
                            update(row, (T)total, deg);
                    } else {
                        update(row, rs[k], deg);
                    }
                }
                pend_rows = 0;
                __syncwarp();
            }
        }

        // row still open at the end of the tile: hand the partial to the tile that completes it
        if (lane == 0 && r_hi < P.n) {
            const int64_t b = P.indptr[r_hi], e = P.indptr[r_hi + 1];
            const int64_t e_lo = item_lo - r_lo, e_hi = item_hi - r_hi;
            const int64_t first_here = b > e_lo ? b : e_lo;
            if (e_hi > first_here) {
                const int64_t t_a = (b + r_hi) / TILE_ITEMS, t_b = (e + r_hi) / TILE_ITEMS;
                const uint32_t expected = (uint32_t)(t_b - t_a + 1);
                double total;
                if (span_arrive(&P.span_acc[t_b], &P.span_cnt[t_b], (double)carry, expected, &total))
                    update(r_hi, (T)total, (int)(e - b));
            }
        }
    }

    if (MODE != MODE_CONV) step_epilogue(P, update, s_red);
}

// Builds the item-space stream from CSR: warp per row.
template <typename V>
__global__ void build_item_stream_kernel(int64_t n, const int32_t *__restrict__ indptr,
                                         const int32_t *__restrict__ indices, const V *__restrict__ values,
                                         int32_t *__restrict__ istream, V *__restrict__ vstream) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n; r += nwarps) {
        const int64_t b = indptr[r], e = indptr[r + 1];
        for (int64_t k = b + lane; k < e; k += 32) {
            istream[k + r] = indices[k];
            if (values) vstream[k + r] = values[k];
        }
        if (lane == 0) {
            istream[e + r] = (int32_t)(-1 - (e - b));
            if (values) vstream[e + r] = (V)0;
        }
    }
}

// Roofline probe (not on the product path): the same evict-first index stream and read-only
// gathers as the fused kernel, nothing else.  Its rate is the ceiling any CSR row-gather can reach
// for this graph on this part; bench.py and DESIGN.md quote the fused kernel against it.
template <typename T>
__global__ void __launch_bounds__(BLOCK, 4) gather_probe_kernel(const int32_t *__restrict__ indices, int64_t nnz,
                                                                const T *__restrict__ z, T *out) {
    T acc = (T)0;
    const int64_t span = (int64_t)BLOCK * IPT;
    for (int64_t base = blockIdx.x * span; base < nnz; base += (int64_t)gridDim.x * span) {
        int32_t cols[IPT];
#pragma unroll
        for (int s = 0; s < IPT; ++s) {
            const int64_t i = base + s * BLOCK + threadIdx.x;
            cols[s] = (i < nnz) ? ld_stream(indices + i) : -1;
        }
#pragma unroll
        for (int s = 0; s < IPT; ++s)
            if (cols[s] >= 0) acc += __ldg(z + cols[s]);
    }
    if (acc == (T)-123456789) out[0] = acc;  // keeps the loads alive
}

static int g_kernel_variant = 4;  // 3 = warp tiles over the item stream, 4 = hsell when the graph carries that form, else 3

template <typename T, bool WEIGHTED, int MODE, bool SYMDEG>
static int launch_tiles(const StepParams &P, cudaStream_t st) {
    static PerDeviceInt ctas_on;
    int &ctas_per_sm = ctas_on.here();
    if (!P.istream || (WEIGHTED && !P.vstream))
        return fail("the item-stream kernel needs pgb_csr.istream%s (pgb_build_item_stream)", WEIGHTED ? "/vstream" : "");
    if (ctas_per_sm == 0) {
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, item_stream_kernel<T, WEIGHTED, MODE, SYMDEG>, BLOCK, 0) !=
                cudaSuccess || v < 1)
            v = 2;
        ctas_per_sm = v;
        // Give shared memory only what the resident CTAs need: the rest of the 228 KB stays L1, which
        // holds the hot end of the gather vector (measured 1.92 -> 1.81 ms on RMAT-24 fp32).
        int pct = -1;
        if (const char *c = getenv("PGB_SMEM_CARVEOUT")) {
            pct = atoi(c);  // experiment knob
        } else {
            cudaFuncAttributes fa;
            if (cudaFuncGetAttributes(&fa, item_stream_kernel<T, WEIGHTED, MODE, SYMDEG>) == cudaSuccess) {
                const size_t need = (fa.sharedSizeBytes + 1024) * (size_t)v;
                pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
                if (pct > 100) pct = 100;
            }
        }
        if (pct >= 0)
            cudaFuncSetAttribute(item_stream_kernel<T, WEIGHTED, MODE, SYMDEG>,
                                 cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    int grid = sm_count() * ctas_per_sm;
    const int need = (int)ceil_div(P.n_tiles, WARPS);
    if (grid > need) grid = need;
    if (grid < 1) return 0;
    item_stream_kernel<T, WEIGHTED, MODE, SYMDEG><<<grid, BLOCK, 0, st>>>(P);
    PGB_LAUNCH_OK("item_stream_kernel");
    return 0;
}

template <int MODE>
static int dispatch(const StepParams &P, int dtype, bool symdeg, cudaStream_t st) {
    const bool weighted = P.values != nullptr;
    if (P.hsell && g_kernel_variant >= 4) {
        if (weighted != (P.hsell->hub_vals != nullptr))
            return fail("the hsell form and the graph disagree about edge values (weighted graph: build the form with values)");
        return hsell_step<MODE>(P, P.hsell, P.partials, dtype, symdeg, st);
    }
    if (dtype == PGB_F32) {
        if (weighted) return launch_tiles<float, true, MODE, false>(P, st);
        if (symdeg) return launch_tiles<float, false, MODE, true>(P, st);
        return launch_tiles<float, false, MODE, false>(P, st);
    } else if (dtype == PGB_F64) {
        if (weighted) return launch_tiles<double, true, MODE, false>(P, st);
        if (symdeg) return launch_tiles<double, false, MODE, true>(P, st);
        return launch_tiles<double, false, MODE, false>(P, st);
    }
    return fail("unknown dtype %d", dtype);
}

static int fill_graph(StepParams &P, const pgb_csr *g) {
    if (!g) return fail("null graph");
    if (g->tile_items != TILE_ITEMS)
        return fail("graph partitioned with tile_items=%d, library uses %d", g->tile_items, TILE_ITEMS);
    if (g->nnz >= (1ll << 31) || g->n >= (1ll << 31)) return fail("graph exceeds the int32 limits of one device");
    if ((int64_t)g->n_tiles != ceil_div(g->n + g->nnz, TILE_ITEMS)) return fail("graph n_tiles inconsistent");
    P.n = g->n;
    P.nnz = g->nnz;
    P.indptr = g->indptr;
    P.indices = g->indices;
    P.values = g->values;
    P.tile_row = g->tile_row;
    P.n_tiles = g->n_tiles;
    P.istream = g->istream;
    P.vstream = g->vstream;
    P.hsell = g->hsell;
    return 0;
}

}  // namespace pgb

using namespace pgb;

extern "C" {

int pgb_tile_items(void) { return TILE_ITEMS; }

int pgb_build_item_stream(int64_t n, int64_t nnz, const int32_t *indptr, const int32_t *indices, int dtype,
                          const void *values, int32_t *istream, void *vstream, void *stream) {
    if (n <= 0) return 0;
    if (n + nnz >= (1ll << 31)) return fail("pgb_build_item_stream: n + nnz exceeds the int32 item space of one device");
    const int grid = stride_grid(n * 32, 256);
    if (!values)
        build_item_stream_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(n, indptr, indices, nullptr, istream, nullptr);
    else if (dtype == PGB_F32)
        build_item_stream_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(n, indptr, indices, (const float *)values,
                                                                              istream, (float *)vstream);
    else if (dtype == PGB_F64)
        build_item_stream_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(n, indptr, indices, (const double *)values,
                                                                               istream, (double *)vstream);
    else
        return fail("pgb_build_item_stream: unknown dtype %d", dtype);
    PGB_LAUNCH_OK("build_item_stream_kernel");
    return 0;
}

int pgb_gather_probe(const pgb_csr *g, int dtype, const void *z, void *scratch, void *stream) {
    if (!g || g->nnz <= 0) return 0;
    const int grid = sm_count() * 4;
    if (dtype == PGB_F32)
        gather_probe_kernel<float><<<grid, BLOCK, 0, as_stream(stream)>>>(g->indices, g->nnz, (const float *)z,
                                                                          (float *)scratch);
    else if (dtype == PGB_F64)
        gather_probe_kernel<double><<<grid, BLOCK, 0, as_stream(stream)>>>(g->indices, g->nnz, (const double *)z,
                                                                           (double *)scratch);
    else
        return fail("pgb_gather_probe: unknown dtype %d", dtype);
    PGB_LAUNCH_OK("gather_probe_kernel");
    return 0;
}

int pgb_set_kernel_variant(int variant) {
    if (variant < 3 || variant > 4) return fail("pgb_set_kernel_variant: %d is not 3 or 4", variant);
    g_kernel_variant = variant;
    return 0;
}

int pgb_spmv(const pgb_csr *g, int dtype, const void *z, const void *rscale, const void *x_for_laplacian,
             const int32_t *out_perm, void *out, pgb_span_ws ws, void *stream) {
    StepParams P;
    memset(&P, 0, sizeof(P));
    if (fill_graph(P, g)) return 1;
    if (P.n == 0) return 0;
    P.zin = z;
    P.zout = out;
    P.rscale = rscale;
    P.xlap = x_for_laplacian;
    P.out_perm = out_perm;
    P.span_acc = ws.acc;
    P.span_cnt = ws.cnt;
    P.partials = ws.partials;
    P.yacc = ws.yacc;
    return dispatch<MODE_CONV>(P, dtype, false, as_stream(stream));
}

int pgb_affine_steps(const pgb_csr *g, int dtype, double alpha, const void *w, const void *sq, const void *c,
                     const void *q, void *zbuf0, void *zbuf1, int64_t out_offset, double *state_f64,
                     int32_t *state_i32, double *err_hist, pgb_span_ws ws, int first_step, int num_launches,
                     int finalize, void *stream) {
    StepParams P;
    memset(&P, 0, sizeof(P));
    if (fill_graph(P, g)) return 1;
    if (P.n == 0) return 0;
    const bool symdeg = (w == nullptr && sq == nullptr);
    if (!symdeg && (!w || !sq)) return fail("pgb_affine_steps: w and sq must both be given or both be NULL");
    if (symdeg && g->values) return fail("pgb_affine_steps: degree-derived scales need an unweighted graph");
    if (first_step < 1) return fail("pgb_affine_steps: first_step must be >= 1");
    P.alpha = alpha;
    P.w = w;
    P.sq = sq;
    P.c = c;
    P.q = q;
    P.out_offset = out_offset;
    P.sf = state_f64;
    P.si = state_i32;
    P.err_hist = err_hist;
    P.span_acc = ws.acc;
    P.span_cnt = ws.cnt;
    P.partials = ws.partials;
    P.yacc = ws.yacc;
    P.finalize = finalize;
    void *buf[2] = {zbuf0, zbuf1};
    for (int j = 0; j < num_launches; ++j) {
        const int k = first_step + j;
        P.zin = buf[(k - 1) & 1];
        P.zout = buf[k & 1];
        P.step = k;
        if (dispatch<MODE_AFFINE>(P, dtype, symdeg, as_stream(stream))) return 1;
    }
    return 0;
}

int pgb_poly_steps(const pgb_csr *g, int dtype, const void *w, const void *sq, const double *coef, void *ranks,
                   void *zbuf0, void *zbuf1, int64_t out_offset, double *state_f64, int32_t *state_i32,
                   double *err_hist, pgb_span_ws ws, int first_step, int num_launches, int finalize, void *stream) {
    StepParams P;
    memset(&P, 0, sizeof(P));
    if (fill_graph(P, g)) return 1;
    if (P.n == 0) return 0;
    const bool symdeg = (w == nullptr && sq == nullptr);
    if (!symdeg && (!w || !sq)) return fail("pgb_poly_steps: w and sq must both be given or both be NULL");
    if (symdeg && g->values) return fail("pgb_poly_steps: degree-derived scales need an unweighted graph");
    if (first_step < 1) return fail("pgb_poly_steps: first_step must be >= 1");
    P.w = w;
    P.sq = sq;
    P.coef = coef;
    P.ranks = ranks;
    P.out_offset = out_offset;
    P.sf = state_f64;
    P.si = state_i32;
    P.err_hist = err_hist;
    P.span_acc = ws.acc;
    P.span_cnt = ws.cnt;
    P.partials = ws.partials;
    P.yacc = ws.yacc;
    P.finalize = finalize;
    void *buf[2] = {zbuf0, zbuf1};
    for (int j = 0; j < num_launches; ++j) {
        const int k = first_step + j;
        P.zin = buf[(k - 1) & 1];
        P.zout = buf[k & 1];
        P.step = k;
        if (dispatch<MODE_POLY>(P, dtype, symdeg, as_stream(stream))) return 1;
    }
    return 0;
}

int pgb_affine_step_peer(const pgb_csr *g, int dtype, double alpha, const void *w, const void *sq, const void *c,
                         const void *q, void *zbuf0, void *zbuf1, int64_t out_offset, double *state_f64,
                         int32_t *state_i32, double *err_hist, pgb_span_ws ws, int step, const pgb_peers *peers,
                         void *stream) {
    StepParams P;
    memset(&P, 0, sizeof(P));
    if (fill_graph(P, g)) return 1;
    if (P.n == 0) return 0;
    if (!peers || peers->n < 1 || peers->n > PGB_MAX_PEERS || peers->rank < 0 || peers->rank >= peers->n)
        return fail("pgb_affine_step_peer: bad peer description");
    const bool symdeg = (w == nullptr && sq == nullptr);
    if (!symdeg && (!w || !sq)) return fail("pgb_affine_step_peer: w and sq must both be given or both be NULL");
    if (symdeg && g->values) return fail("pgb_affine_step_peer: degree-derived scales need an unweighted graph");
    if (step < 1) return fail("pgb_affine_step_peer: step must be >= 1");
    P.alpha = alpha;
    P.w = w;
    P.sq = sq;
    P.c = c;
    P.q = q;
    P.out_offset = out_offset;
    P.sf = state_f64;
    P.si = state_i32;
    P.err_hist = err_hist;
    P.span_acc = ws.acc;
    P.span_cnt = ws.cnt;
    P.partials = ws.partials;
    P.yacc = ws.yacc;
    P.finalize = 0;
    void *buf[2] = {zbuf0, zbuf1};
    P.zin = buf[(step - 1) & 1];
    P.zout = buf[step & 1];
    P.step = step;
    P.n_peers = peers->n;
    P.peer_rank = peers->rank;
    for (int r = 0; r < peers->n; ++r) {
        P.peer_zout[r] = (step & 1) ? peers->zbuf1[r] : peers->zbuf0[r];
        P.peer_acc[r] = peers->acc[r];
        if (!P.peer_zout[r] || !P.peer_acc[r]) return fail("pgb_affine_step_peer: peer %d has no buffer", r);
    }
    P.mc_zout = (step & 1) ? peers->mc_zbuf1 : peers->mc_zbuf0;
    P.peer_mask = peers->row_mask;
    return dispatch<MODE_AFFINE>(P, dtype, symdeg, as_stream(stream));
}

// pgb_poly_steps for ONE step with the exchange fused into it (see pgb_affine_step_peer).
int pgb_poly_step_peer(const pgb_csr *g, int dtype, const void *w, const void *sq, const double *coef, void *ranks,
                       void *zbuf0, void *zbuf1, int64_t out_offset, double *state_f64, int32_t *state_i32,
                       double *err_hist, pgb_span_ws ws, int step, const pgb_peers *peers, void *stream) {
    StepParams P;
    memset(&P, 0, sizeof(P));
    if (fill_graph(P, g)) return 1;
    if (P.n == 0) return 0;
    if (!peers || peers->n < 1 || peers->n > PGB_MAX_PEERS || peers->rank < 0 || peers->rank >= peers->n)
        return fail("pgb_poly_step_peer: bad peer description");
    const bool symdeg = (w == nullptr && sq == nullptr);
    if (!symdeg && (!w || !sq)) return fail("pgb_poly_step_peer: w and sq must both be given or both be NULL");
    if (symdeg && g->values) return fail("pgb_poly_step_peer: degree-derived scales need an unweighted graph");
    if (step < 1) return fail("pgb_poly_step_peer: step must be >= 1");
    P.w = w;
    P.sq = sq;
    P.coef = coef;
    P.ranks = ranks;
    P.out_offset = out_offset;
    P.sf = state_f64;
    P.si = state_i32;
    P.err_hist = err_hist;
    P.span_acc = ws.acc;
    P.span_cnt = ws.cnt;
    P.partials = ws.partials;
    P.yacc = ws.yacc;
    P.finalize = 0;
    void *buf[2] = {zbuf0, zbuf1};
    P.zin = buf[(step - 1) & 1];
    P.zout = buf[step & 1];
    P.step = step;
    P.n_peers = peers->n;
    P.peer_rank = peers->rank;
    for (int r = 0; r < peers->n; ++r) {
        P.peer_zout[r] = (step & 1) ? peers->zbuf1[r] : peers->zbuf0[r];
        P.peer_acc[r] = peers->acc[r];
        if (!P.peer_zout[r] || !P.peer_acc[r]) return fail("pgb_poly_step_peer: peer %d has no buffer", r);
    }
    P.mc_zout = (step & 1) ? peers->mc_zbuf1 : peers->mc_zbuf0;
    P.peer_mask = peers->row_mask;
    return dispatch<MODE_POLY>(P, dtype, symdeg, as_stream(stream));
}

__global__ void state_finalize_peer_kernel(double *sf, int32_t *si, double *err_hist, const double *slots, int n) {
    if (si[PGB_SI_STOP] != PGB_RUNNING) return;
    double t = 0.0, e = 0.0;
    const bool is_max = si[PGB_SI_ERR_MODE] == PGB_ERR_MAX;
    for (int r = 0; r < n; ++r) {   // rank order: the same sum on every rank
        t += ((const volatile double *)slots)[2 * r];
        const double er = ((const volatile double *)slots)[2 * r + 1];
        e = is_max ? fmax(e, er) : e + er;
    }
    sf[PGB_SF_TACC] = t;
    sf[PGB_SF_EACC] = e;
    finalize_state(sf, si, err_hist);
}

int pgb_state_finalize_peer(double *state_f64, int32_t *state_i32, double *err_hist, const double *acc_slots,
                            int32_t n, void *stream) {
    if (!acc_slots || n < 1) return fail("pgb_state_finalize_peer: no slots");
    state_finalize_peer_kernel<<<1, 1, 0, as_stream(stream)>>>(state_f64, state_i32, err_hist, acc_slots, n);
    PGB_LAUNCH_OK("state_finalize_peer_kernel");
    return 0;
}

int pgb_state_finalize(double *state_f64, int32_t *state_i32, double *err_hist, void *stream) {
    state_finalize_kernel<<<1, 1, 0, as_stream(stream)>>>(state_f64, state_i32, err_hist);
    PGB_LAUNCH_OK("state_finalize_kernel");
    return 0;
}

}  // extern "C"
